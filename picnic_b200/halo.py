"""Multi-box exchanges of the particle engine: one process per GPU, one or several Chombo boxes per process.

What the reference does with MPI inside Chombo -- `LevelData::exchange` with an add op on the
ghost cells of J (PicSpeciesInterface::finalizeSettingJ, PicSpeciesInterface.cpp:766-772) and
`ParticleData::remapOutcast` (ParticleDataI.H:405-547) -- is done here with point-to-point
messages between the boxes of the square equal-box decomposition (System.cpp:169-245), on DEVICE
buffers over NCCL/NVLink (`torch.distributed`; gloo on CPU in the tests).  The arithmetic on the
library's arrays stays in the CUDA library (pgpu_fab_pack_d / pgpu_fab_unpack_d /
pgpu_species_mark_leavers / _pack_leavers_d / _append_d); this module only computes index boxes
and moves buffers.  Collisions and moments are cell-local and need no exchange.

The add-exchange runs direction by direction over the full transverse extent of the arrays
(ghosts included), so corners need no diagonal messages: after the pass in direction d every
index that two boxes share in d holds the sum of both, and the later directions carry those
sums on.  Directions in which the box spans the whole periodic domain are folded locally by
`pgpu_current_finalize`.
"""
import numpy as np

STAG_J = {1: [(0,), (1,), (1,)], 2: [(0, 1), (1, 0), (1, 1)]}   # Jx, Jy, Jz centring (1 = nodal)


class BoxLayout:
    """px x py equal boxes of nbox cells tiling a domain of ncell cells; rank = bi + bj*px."""

    def __init__(self, D, ncell, nbox, nghost, periodic):
        self.D, self.ncell, self.nbox, self.nghost = D, tuple(ncell), tuple(nbox), nghost
        self.periodic = tuple(int(p) for p in periodic)
        self.nb = tuple(nc // nb for nc, nb in zip(ncell, nbox))
        for nc, nb in zip(ncell, nbox):
            assert nc % nb == 0, "boxes must tile the domain"
        for d in range(D):
            assert self.nb[d] == 1 or nbox[d] >= 2 * nghost + 1, "box narrower than its ghost overlap"
        self.world = int(np.prod(self.nb))

    def coords(self, rank):
        return (rank % self.nb[0],) if self.D == 1 else (rank % self.nb[0], rank // self.nb[0])

    def rank_of(self, c):
        return c[0] if self.D == 1 else c[0] + c[1] * self.nb[0]

    def box(self, rank):
        c = self.coords(rank)
        lo = tuple(ci * nb for ci, nb in zip(c, self.nbox))
        hi = tuple(l + nb - 1 for l, nb in zip(lo, self.nbox))
        return lo, hi

    def neighbor(self, rank, d, side):
        """Rank of the neighbour box in direction d (side -1 / +1), or None at a wall."""
        c = list(self.coords(rank))
        c[d] += side
        if c[d] < 0 or c[d] >= self.nb[d]:
            if not self.periodic[d]:
                return None
            c[d] %= self.nb[d]
        return self.rank_of(c)

    def neighbor_code(self, rank, code):
        """Rank that owns direction code (d0+1) + 3 (d1+1) of `rank` (migration)."""
        off = (code % 3 - 1, code // 3 - 1)
        c = list(self.coords(rank))
        for d in range(self.D):
            c[d] += off[d]
            if c[d] < 0 or c[d] >= self.nb[d]:
                if not self.periodic[d]:
                    return None
                c[d] %= self.nb[d]
        return self.rank_of(c)

    def array_bounds(self, rank, stag):
        lo, hi = self.box(rank)
        g = self.nghost
        return (tuple(l - g for l in lo), tuple(h + g + s for h, s in zip(hi, stag)))

    def overlap(self, rank, stag, d, side):
        """Index box (in this rank's global indices) shared with the neighbour in direction d:
        2*nghost + stag[d] layers around the common face, full extent elsewhere."""
        alo, ahi = self.array_bounds(rank, stag)
        lo, hi = list(alo), list(ahi)
        blo, bhi = self.box(rank)
        g = self.nghost
        if side > 0:
            lo[d], hi[d] = bhi[d] + 1 - g, bhi[d] + g + stag[d]
        else:
            lo[d], hi[d] = blo[d] - g, blo[d] - 1 + g + stag[d]
        return tuple(lo), tuple(hi)


# ------------------------------------------------------------------------------------------------
# communicators
# ------------------------------------------------------------------------------------------------
class DistComm:
    """torch.distributed point-to-point (NCCL on GPUs, gloo on CPU)."""

    def __init__(self, rank, world, stream=None):
        """stream: the torch CUDA stream the library runs on (pgpu_set_stream).  Collectives are
        enqueued relative to it, so pack -> send and recv -> unpack order on the device without
        a host synchronisation."""
        import contextlib
        import torch
        import torch.distributed as dist
        self.dist, self.rank, self.world, self.torch = dist, rank, world, torch
        self.ctx = (lambda: torch.cuda.stream(stream)) if stream is not None else contextlib.nullcontext
        self.reqs = []

    def post(self, sends, recvs):
        ops = [self.dist.P2POp(self.dist.isend, t, peer, tag=tag) for (peer, tag, t) in sends]
        ops += [self.dist.P2POp(self.dist.irecv, t, peer, tag=tag) for (peer, tag, t) in recvs]
        with self.ctx():
            self.reqs = self.dist.batch_isend_irecv(ops) if ops else []

    def wait(self):
        with self.ctx():
            for r in self.reqs:
                r.wait()
        self.reqs = []

    def all_gather(self, t):
        """Host copies of every rank's tensor (the read-back happens on the communicator's stream)."""
        with self.ctx():
            t = t.clone()
            out = [self.torch.empty_like(t) for _ in range(self.world)]
            self.dist.all_gather(out, t)
            return [o.cpu() for o in out]


class LoopComm:
    """All boxes in ONE process (tests, and several boxes per GPU): a mailbox keyed by
    (source, destination, tag).  Call post() for every box before wait() for any."""

    def __init__(self, world):
        self.world, self.box, self.gathered = world, {}, {}

    def view(self, rank):
        return _LoopView(self, rank)


class _LoopView:
    def __init__(self, hub, rank):
        self.hub, self.rank, self.world = hub, rank, hub.world
        self.recvs = []

    def post(self, sends, recvs):
        for (peer, tag, t) in sends:
            self.hub.box[(self.rank, peer, tag)] = t.clone()
        self.recvs = recvs

    def wait(self):
        for (peer, tag, t) in self.recvs:
            t.copy_(self.hub.box.pop((peer, self.rank, tag)))
        self.recvs = []

    def all_gather_post(self, t):
        self.hub.gathered[self.rank] = t.clone()

    def all_gather_collect(self):
        return [self.hub.gathered[r] for r in range(self.world)]


# ------------------------------------------------------------------------------------------------
# ghost add-exchange of the total current
# ------------------------------------------------------------------------------------------------
class CapiGridBackend:
    """Device buffers <-> the library's total-J arrays through the C ABI."""

    def __init__(self, grid, device, on_torch_stream=False):
        """on_torch_stream: the library runs on the torch stream the communicator enqueues on
        (no host synchronisation between pack and send)."""
        import torch
        from . import capi
        self.torch, self.capi, self.grid, self.device = torch, capi, grid, device
        self.on_torch_stream = on_torch_stream

    def new_buffer(self, count):
        return self.torch.empty(count, dtype=self.torch.float64, device=self.device)

    def pack(self, comp, lo, hi, buf):
        c = self.capi
        c.check(c.load().pgpu_fab_pack_d(self.grid.h, 0, comp, c._i2(lo), c._i2(hi), buf.data_ptr()))

    def unpack_add(self, comp, lo, hi, buf):
        c = self.capi
        c.check(c.load().pgpu_fab_unpack_d(self.grid.h, 0, comp, c._i2(lo), c._i2(hi), buf.data_ptr(), 1))

    def sync(self):
        if not self.on_torch_stream:   # the library runs on its own stream; collectives on torch's
            self.capi.check(self.capi.load().pgpu_synchronize())


class HaloExchange:
    """Ghost ADD-exchange of Jx, Jy, Jz between neighbouring boxes, direction by direction.  The three
    components that go to one neighbour travel in ONE message (one send and one receive per side and
    direction: the cost of these KB-scale exchanges is per message, not per byte)."""

    def __init__(self, layout, rank, comm, backend):
        self.layout, self.rank, self.comm, self.be = layout, rank, comm, backend
        self.plan = []   # per direction: list of (peer, send_tag, recv_tag, parts, sendbuf, recvbuf)
        D = layout.D     # parts = [(comp, lo, hi, offset, count)] inside the side's buffer
        for d in range(D):
            if layout.nb[d] == 1:
                continue          # spans the domain: folded locally (pgpu_current_finalize)
            msgs = []
            for side in (-1, +1):
                peer = layout.neighbor(rank, d, side)
                if peer is None:
                    continue
                parts, off = [], 0
                for comp, stag in enumerate(STAG_J[D]):
                    lo, hi = layout.overlap(rank, stag, d, side)
                    count = int(np.prod([h - l + 1 for l, h in zip(lo, hi)]))
                    parts.append((comp, lo, hi, off, count))
                    off += count
                # my +side message is the peer's -side message: tag by the receiver's side
                send_tag = 16 * d + 4 * (0 if side > 0 else 1)
                recv_tag = 16 * d + 4 * (1 if side > 0 else 0)
                msgs.append((peer, send_tag, recv_tag, parts, self.be.new_buffer(off), self.be.new_buffer(off)))
            self.plan.append(msgs)
        self.bytes_per_exchange = sum(2 * 8 * m[4].numel() for msgs in self.plan for m in msgs)

    def n_phases(self):
        return len(self.plan)

    def begin(self, phase):
        msgs = self.plan[phase]
        for (_, _, _, parts, sb, _) in msgs:
            for (comp, lo, hi, off, count) in parts:
                self.be.pack(comp, lo, hi, sb[off:off + count])
        self.be.sync()
        self.comm.post([(m[0], m[1], m[4]) for m in msgs], [(m[0], m[2], m[5]) for m in msgs])

    def end(self, phase):
        self.comm.wait()
        for (_, _, _, parts, _, rb) in self.plan[phase]:
            for (comp, lo, hi, off, count) in parts:
                self.be.unpack_add(comp, lo, hi, rb[off:off + count])

    def add_exchange(self):
        """One process per box: the whole exchange (every rank calls this collectively)."""
        for ph in range(self.n_phases()):
            self.begin(ph)
            self.end(ph)


class PeerHaloExchange:
    """The same direction-by-direction plan as HaloExchange, executed by the library over peer memory
    (pgpu_halo_*, csrc/pgpu_halo_p2p.cu): per direction one kernel stores this box's overlap strips into
    the neighbours' inboxes over NVLink and stamps their arrival flags, a second one waits for this box's
    own flags and adds what arrived.  No NCCL call and no host synchronisation on the data path; the
    only collective is the one-off exchange of the CUDA IPC handles of the inboxes (connect_*)."""

    def __init__(self, layout, rank, grid, rho_stag=None):
        """rho_stag = None: the plan for the three J components; a centring tuple: the plan for the grid's resident
        charge-density array of that centring (pgpu_halo_create_rho, PicChargedSpecies.cpp:3083/:3120/:3141)."""
        from . import capi
        self.capi, self.layout, self.rank, self.grid = capi, layout, rank, grid
        D = layout.D
        stags = STAG_J[D] if rho_stag is None else [tuple(int(v) for v in rho_stag[:D])]
        self.msgs = []          # (phase, side, peer)
        recs = []
        phase = 0
        for d in range(D):
            if layout.nb[d] == 1:
                continue
            for side in (-1, +1):
                peer = layout.neighbor(rank, d, side)
                if peer is None:
                    continue
                m = capi.HaloMsg()
                m.phase, m.recv_area = phase, len(recs)
                for comp, stag in enumerate(stags):
                    lo, hi = layout.overlap(rank, stag, d, side)
                    for k in range(D):
                        m.lo[comp][k], m.hi[comp][k] = lo[k], hi[k]
                recs.append(m)
                self.msgs.append((phase, side, peer))
            phase += 1
        arr = (capi.HaloMsg * max(len(recs), 1))(*recs)
        self.h = capi.C.c_void_p()
        if rho_stag is None:
            capi.check(capi.load().pgpu_halo_create(grid.h, len(recs), arr, capi.C.byref(self.h)))
        else:
            st = (capi.C.c_int * 2)(int(rho_stag[0]), int(rho_stag[1]) if D == 2 else 0)
            capi.check(capi.load().pgpu_halo_create_rho(grid.h, st, len(recs), arr, capi.C.byref(self.h)))
        self.nphase = capi.load().pgpu_halo_phases(self.h)
        self.bytes_per_exchange = 0
        self.areas = {}         # (phase, side) -> (area index, offset in doubles)
        for i, (ph, side, _) in enumerate(self.msgs):
            off, cnt = capi.C.c_long(), capi.C.c_long()
            capi.check(capi.load().pgpu_halo_area_offset(self.h, i, capi.C.byref(off), capi.C.byref(cnt)))
            self.areas[(ph, side)] = (i, off.value)
            self.bytes_per_exchange += 2 * 8 * cnt.value

    # ---- wiring -------------------------------------------------------------------------------
    def inbox_pointer(self):
        p, n = self.capi.C.c_void_p(), self.capi.C.c_size_t()
        self.capi.check(self.capi.load().pgpu_halo_inbox(self.h, self.capi.C.byref(p), self.capi.C.byref(n)))
        return p.value

    def _connect(self, pointer_of, areas_of):
        lib, capi = self.capi.load(), self.capi
        for i, (ph, side, peer) in enumerate(self.msgs):
            area, off = areas_of(peer)[(ph, -side)]       # my +side message is the peer's -side arrival
            capi.check(lib.pgpu_halo_connect(self.h, i, capi.C.c_void_p(pointer_of(peer)), area, off))

    @staticmethod
    def connect_local(exchanges):
        """All boxes in this process (tests; several boxes per GPU): plain device pointers."""
        by_rank = {e.rank: e for e in exchanges}
        for e in exchanges:
            e._connect(lambda r: by_rank[r].inbox_pointer(), lambda r: by_rank[r].areas)

    def connect_ipc(self, comm):
        """One process per box: all-gather the 64-byte CUDA IPC handle and the area table of every inbox,
        open the neighbours' handles."""
        import torch
        capi, lib = self.capi, self.capi.load()
        hbuf = (capi.C.c_ubyte * 64)()
        capi.check(lib.pgpu_halo_ipc_handle(self.h, hbuf))
        rec = np.zeros(8 + 1 + 4 * 16, dtype=np.int64)
        rec[:8] = np.frombuffer(bytes(hbuf), dtype=np.int64)
        rec[8] = len(self.msgs)
        for i, ((ph, side), (area, off)) in enumerate(sorted(self.areas.items())):
            rec[9 + 4 * i: 13 + 4 * i] = (ph, side, area, off)
        device = torch.device("cuda", torch.cuda.current_device()) if comm.dist.get_backend() == "nccl" else "cpu"
        allr = [x.numpy() for x in comm.all_gather(torch.as_tensor(rec).to(device))]
        opened = {}

        def pointer_of(r):
            if r not in opened:
                hb = (capi.C.c_ubyte * 64).from_buffer_copy(allr[r][:8].tobytes())
                p = capi.C.c_void_p()
                if r == self.rank:
                    opened[r] = self.inbox_pointer()
                else:
                    capi.check(lib.pgpu_halo_ipc_open(self.h, hb, capi.C.byref(p)))
                    opened[r] = p.value
            return opened[r]

        def areas_of(r):
            n = int(allr[r][8])
            return {(int(a[0]), int(a[1])): (int(a[2]), int(a[3])) for a in allr[r][9:9 + 4 * n].reshape(n, 4)}

        self._connect(pointer_of, areas_of)

    @staticmethod
    def connect_mixed(exchanges, comm, boxes_per_rank):
        """Several boxes per process, several processes (C5: sixteen boxes, two per GPU): boxes of this process are wired
        with plain device pointers, the others through their CUDA IPC handles.  `exchanges` = this process's
        PeerHaloExchange objects in box order; box b lives in process b // boxes_per_rank."""
        import torch
        e0 = exchanges[0]
        capi, lib = e0.capi, e0.capi.load()
        B = boxes_per_rank
        assert len(exchanges) == B
        nrec = 8 + 1 + 4 * 16
        rec = np.zeros((B, nrec), dtype=np.int64)
        for k, e in enumerate(exchanges):
            hbuf = (capi.C.c_ubyte * 64)()
            capi.check(lib.pgpu_halo_ipc_handle(e.h, hbuf))
            rec[k, :8] = np.frombuffer(bytes(hbuf), dtype=np.int64)
            rec[k, 8] = len(e.msgs)
            for i, ((ph, side), (area, off)) in enumerate(sorted(e.areas.items())):
                rec[k, 9 + 4 * i: 13 + 4 * i] = (ph, side, area, off)
        device = torch.device("cuda", torch.cuda.current_device()) if comm.dist.get_backend() == "nccl" else "cpu"
        allr = [x.numpy() for x in comm.all_gather(torch.as_tensor(rec).to(device))]     # [process][box, nrec]
        local = {e.rank: e for e in exchanges}
        opened = {}

        def areas_of(b):
            r = allr[b // B][b % B]
            n = int(r[8])
            return {(int(a[0]), int(a[1])): (int(a[2]), int(a[3])) for a in r[9:9 + 4 * n].reshape(n, 4)}

        for e in exchanges:
            def pointer_of(b, e=e):
                if b in local:
                    return local[b].inbox_pointer()
                if b not in opened:
                    hb = (capi.C.c_ubyte * 64).from_buffer_copy(allr[b // B][b % B][:8].tobytes())
                    ptr = capi.C.c_void_p()
                    capi.check(lib.pgpu_halo_ipc_open(e.h, hb, capi.C.byref(ptr)))
                    opened[b] = ptr.value
                return opened[b]
            e._connect(pointer_of, areas_of)

    # ---- the exchange -------------------------------------------------------------------------
    def begin(self):
        self.capi.check(self.capi.load().pgpu_halo_begin(self.h))

    def send(self, phase):
        self.capi.check(self.capi.load().pgpu_halo_send(self.h, phase))

    def recv_add(self, phase):
        self.capi.check(self.capi.load().pgpu_halo_recv_add(self.h, phase))

    def add_exchange(self):
        """One process per box: the whole exchange (every rank calls this collectively)."""
        self.begin()
        for ph in range(self.nphase):
            self.send(ph)
            self.recv_add(ph)

    def destroy(self):
        if self.h:
            self.capi.check(self.capi.load().pgpu_halo_destroy(self.h))
            self.h = None


# ------------------------------------------------------------------------------------------------
# particle migration
# ------------------------------------------------------------------------------------------------
class CapiSpeciesBackend:
    def __init__(self, species, device, on_torch_stream=False):
        import torch
        from . import capi
        self.torch, self.capi, self.sp, self.device = torch, capi, species, device
        self.nw = capi.load().pgpu_wire_doubles(species.grid.h)
        self.on_torch_stream = on_torch_stream

    def mark_leavers(self):
        import ctypes
        cnt = (ctypes.c_long * 10)()
        self.capi.check(self.capi.load().pgpu_species_mark_leavers(self.sp.h, cnt))
        return np.array(list(cnt), dtype=np.int64)

    def mark_leavers_device(self, counts_d):
        """No host synchronisation: counts land in the device int64[10] tensor counts_d."""
        self.capi.check(self.capi.load().pgpu_species_mark_leavers_d(self.sp.h, counts_d.data_ptr()))

    def set_leaver_counts(self, counts):
        import ctypes
        cnt = (ctypes.c_long * 10)(*[int(v) for v in counts])
        self.capi.check(self.capi.load().pgpu_species_set_leaver_counts(self.sp.h, cnt))

    def new_buffer(self, nrec):
        return self.torch.empty(max(nrec, 1) * self.nw, dtype=self.torch.float64, device=self.device)

    def pack_leavers(self, buf):
        self.capi.check(self.capi.load().pgpu_species_pack_leavers_d(self.sp.h, buf.data_ptr()))

    def append(self, nrec, buf):
        self.capi.check(self.capi.load().pgpu_species_append_d(self.sp.h, int(nrec), buf.data_ptr()))

    def sync(self):
        if not self.on_torch_stream:
            self.capi.check(self.capi.load().pgpu_synchronize())


class Migration:
    """Outgoing particles to the owning neighbour box, once per step after applyBCs."""

    def __init__(self, layout, rank, comm, backend):
        self.layout, self.rank, self.comm, self.be = layout, rank, comm, backend
        self.sendbuf = None
        self.counts = None
        self.lost = 0

    def begin_counts(self):
        import torch
        c = self.be.mark_leavers()
        self.lost = int(c[9])
        self.counts = c[:9].copy()
        self.counts[4] = 0
        return torch.as_tensor(self.counts)

    def counts_on_device(self, t):
        return t.to(self.be.device)

    def begin_payload(self, all_counts):
        """all_counts[r] = the 9 per-direction leaver counts of rank r."""
        lay, me, nw = self.layout, self.rank, self.be.nw
        total = int(self.counts.sum())
        self.sendbuf = self.be.new_buffer(total)
        self.be.pack_leavers(self.sendbuf)
        self.be.sync()
        sends, recvs, self.arrivals = [], [], []
        off = 0
        for code in range(9):
            n = int(self.counts[code])
            if n == 0:
                continue
            peer = lay.neighbor_code(me, code)
            assert peer is not None and peer != me, "leaver without an owner box"
            sends.append((peer, code, self.sendbuf[off * nw:(off + n) * nw]))
            off += n
        for r in range(lay.world):
            if r == me:
                continue
            cr = [int(v) for v in all_counts[r]]
            for code in range(9):
                if cr[code] and lay.neighbor_code(r, code) == me:
                    buf = self.be.new_buffer(cr[code])
                    recvs.append((r, code, buf[:cr[code] * nw]))
                    self.arrivals.append((cr[code], buf))
        self.comm.post(sends, recvs)

    def end(self):
        self.comm.wait()
        n_in = 0
        for (n, buf) in self.arrivals:
            self.be.append(n, buf)
            n_in += n
        self.be.sync()
        self.sendbuf, self.arrivals = None, []
        return n_in

    def use_counts(self, c10):
        """Counts of this box obtained elsewhere (migrate_all's single device all-gather)."""
        c10 = np.asarray(c10, dtype=np.int64)
        self.lost = int(c10[9])
        self.counts = c10[:9].copy()
        self.counts[4] = 0
        full = np.zeros(10, dtype=np.int64)
        full[:9] = self.counts
        full[9] = self.lost           # removed from the species, not sent anywhere
        self.be.set_leaver_counts(full)

    def migrate(self):
        """One process per box: the whole migration (collective)."""
        t = self.counts_on_device(self.begin_counts())
        allc = [x.numpy() for x in self.comm.all_gather(t)]
        self.begin_payload(allc)
        return self.end()


def migrate_all(migrations):
    """Migration of several species of one box with ONE host synchronisation: every species is
    marked on the device, the count vectors of all species and boxes are all-gathered on the
    device and read back once, then the payloads move species by species."""
    if not migrations:
        return 0
    m0 = migrations[0]
    torch = m0.be.torch
    dev_counts = torch.zeros((len(migrations), 10), dtype=torch.int64, device=m0.be.device)
    for k, m in enumerate(migrations):
        m.be.mark_leavers_device(dev_counts[k])
    allc = [x.numpy() for x in m0.comm.all_gather(dev_counts)]        # [rank][species, 10]
    n_in = 0
    for k, m in enumerate(migrations):
        m.use_counts(allc[m.rank][k])
        m.begin_payload([allc[r][k][:9] for r in range(m.layout.world)])
        n_in += m.end()
    return n_in


class PeerMigration:
    """Migration of one species of one box over peer memory (pgpu_migrator_*, csrc/pgpu_exchange.cu):
    leavers are stored straight into the owning neighbours' inboxes by the sending kernel, counts stay on
    the device, the receiving kernel appends; the only host synchronisation is finish().  The CUDA IPC
    handles of the inboxes are exchanged once (connect_ipc) -- nothing else goes through a message layer."""

    def __init__(self, layout, rank, species, capacity):
        from . import capi
        self.capi, self.layout, self.rank, self.sp = capi, layout, rank, species
        self.h = capi.C.c_void_p()
        capi.check(capi.load().pgpu_migrator_create(species.h, int(capacity), capi.C.byref(self.h)))
        self.capacity = int(capacity)
        self.lost = 0

    def neighbours(self):
        """{direction code: rank} of the boxes that can own a leaver of this one."""
        out = {}
        for code in range(9):
            if code == 4:
                continue
            peer = self.layout.neighbor_code(self.rank, code)
            if peer is not None and peer != self.rank:
                out[code] = peer
        return out

    def inbox_pointer(self):
        p, n = self.capi.C.c_void_p(), self.capi.C.c_size_t()
        self.capi.check(self.capi.load().pgpu_migrator_inbox(self.h, self.capi.C.byref(p), self.capi.C.byref(n)))
        return p.value

    @staticmethod
    def connect_local(migrations):
        by_rank = {m.rank: m for m in migrations}
        for m in migrations:
            for code, peer in m.neighbours().items():
                m.capi.check(m.capi.load().pgpu_migrator_connect(m.h, code, m.capi.C.c_void_p(by_rank[peer].inbox_pointer())))

    def connect_ipc(self, comm):
        import torch
        capi, lib = self.capi, self.capi.load()
        hbuf = (capi.C.c_ubyte * 64)()
        capi.check(lib.pgpu_migrator_ipc_handle(self.h, hbuf))
        rec = np.zeros(9, dtype=np.int64)
        rec[:8] = np.frombuffer(bytes(hbuf), dtype=np.int64)
        rec[8] = self.capacity
        device = torch.device("cuda", torch.cuda.current_device()) if comm.dist.get_backend() == "nccl" else "cpu"
        allr = [x.numpy() for x in comm.all_gather(torch.as_tensor(rec).to(device))]
        opened = {}
        for code, peer in self.neighbours().items():
            assert int(allr[peer][8]) == self.capacity, "migration inboxes must have one capacity"
            if peer not in opened:
                hb = (capi.C.c_ubyte * 64).from_buffer_copy(allr[peer][:8].tobytes())
                p = capi.C.c_void_p()
                capi.check(lib.pgpu_migrator_ipc_open(self.h, hb, capi.C.byref(p)))
                opened[peer] = p.value
            capi.check(lib.pgpu_migrator_connect(self.h, code, capi.C.c_void_p(opened[peer])))

    @staticmethod
    def connect_mixed(migrations, comm, boxes_per_rank):
        """Several boxes per process, several processes: `migrations` = the PeerMigration objects of ONE species for
        this process's boxes, in box order (box b lives in process b // boxes_per_rank)."""
        import torch
        m0 = migrations[0]
        capi, lib = m0.capi, m0.capi.load()
        B = boxes_per_rank
        assert len(migrations) == B
        rec = np.zeros((B, 9), dtype=np.int64)
        for k, m in enumerate(migrations):
            hbuf = (capi.C.c_ubyte * 64)()
            capi.check(lib.pgpu_migrator_ipc_handle(m.h, hbuf))
            rec[k, :8] = np.frombuffer(bytes(hbuf), dtype=np.int64)
            rec[k, 8] = m.capacity
        device = torch.device("cuda", torch.cuda.current_device()) if comm.dist.get_backend() == "nccl" else "cpu"
        allr = [x.numpy() for x in comm.all_gather(torch.as_tensor(rec).to(device))]
        local = {m.rank: m for m in migrations}
        opened = {}
        for m in migrations:
            for code, peer in m.neighbours().items():
                assert int(allr[peer // B][peer % B][8]) == m.capacity, "migration inboxes must have one capacity"
                if peer in local:
                    ptr = local[peer].inbox_pointer()
                else:
                    if peer not in opened:
                        hb = (capi.C.c_ubyte * 64).from_buffer_copy(allr[peer // B][peer % B][:8].tobytes())
                        q = capi.C.c_void_p()
                        capi.check(lib.pgpu_migrator_ipc_open(m.h, hb, capi.C.byref(q)))
                        opened[peer] = q.value
                    ptr = opened[peer]
                capi.check(lib.pgpu_migrator_connect(m.h, code, capi.C.c_void_p(ptr)))

    def send(self):
        self.capi.check(self.capi.load().pgpu_migrate_send(self.h))

    def recv(self):
        self.capi.check(self.capi.load().pgpu_migrate_recv(self.h))

    def finish(self):
        C = self.capi.C
        a, l, x = C.c_long(), C.c_long(), C.c_long()
        self.capi.check(self.capi.load().pgpu_migrate_finish(self.h, C.byref(a), C.byref(l), C.byref(x)))
        self.lost = x.value
        return a.value

    def destroy(self):
        if self.h:
            self.capi.check(self.capi.load().pgpu_migrator_destroy(self.h))
            self.h = None


def migrate_all_peer(migrations):
    """All species of this box (or all boxes of this process): send everything, then receive everything,
    then ONE wait."""
    for m in migrations:
        m.send()
    for m in migrations:
        m.recv()
    return sum(m.finish() for m in migrations)
