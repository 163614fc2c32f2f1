"""Device side of the multi-box exchanges on one GPU: two / four boxes live in one process and
exchange through the in-process mailbox (picnic_b200.halo.LoopComm), so that the pack /
unpack-add / mark / pack-leavers / append kernels run on real device buffers.  The result must
equal the single-box run of the same particles (the gloo test covers the process plumbing)."""
import numpy as np
import pytest

from picnic_b200 import halo

pytestmark = pytest.mark.gpu

NCELL, NG = (32, 16), 2
DX, XMIN = (0.25, 0.5), (0.0, -1.0)


def _particles(seed, n=20000):
    rng = np.random.default_rng(seed)
    L = np.array([nc * h for nc, h in zip(NCELL, DX)])
    xo = np.array(XMIN)[:, None] + rng.random((2, n)) * L[:, None]
    x = xo + (rng.random((2, n)) - 0.5) * np.array(DX)[:, None] * 0.8
    v = rng.standard_normal((3, n)) * 0.05
    w = rng.random(n) + 0.5
    return np.ascontiguousarray(x), np.ascontiguousarray(xo), v, w


def _species(pgpu, grid, x, xo, v, w, ids):
    sp = pgpu.Species(grid, 1.0, -1.0, 1.0, 1.0, interp_N=1, interp_J=3, interp_E=3)
    sp.upload(x, v, w, xold=xo, vold=v, ids=ids)
    return sp


@pytest.mark.parametrize("nbox,peer", [((16, 16), False), ((16, 8), False), ((16, 16), True), ((16, 8), True),
                                       ((8, 16), True)])
def test_add_exchange_matches_single_box(pgpu, nbox, peer):
    """peer=False: pack / mailbox / unpack-add (the NCCL route's kernels); peer=True: the peer-memory
    kernels (pgpu_halo_send / pgpu_halo_recv_add), three exchanges in a row so that both parity slots of
    the inboxes and the sequence numbers are exercised."""
    import torch
    lay = halo.BoxLayout(2, NCELL, nbox, NG, (1, 1))
    x, xo, v, w = _particles(2)
    ids = np.arange(w.size, dtype=np.uint64)
    # single box spanning the domain
    g1 = pgpu.Grid(2, NCELL, XMIN, DX, NG, (1, 1))
    s1 = _species(pgpu, g1, x, xo, v, w, ids)
    s1.set_current_density(1.0)
    g1.current_zero(); g1.current_add(s1); g1.current_finalize()
    Jg = [(g1.field_bounds(c), g1.current_get(c)) for c in range(3)]
    s1.destroy(); g1.destroy()
    # the boxes of the decomposition
    own = sum(np.floor((xo[d] - XMIN[d]) / (DX[d] * nbox[d])).astype(int) * (1 if d == 0 else lay.nb[0])
              for d in range(2))
    hub = halo.LoopComm(lay.world)
    grids, sps, hxs = [], [], []
    for r in range(lay.world):
        lo, hi = lay.box(r)
        g = pgpu.Grid(2, NCELL, XMIN, DX, NG, (1, 1), box_lo=lo, box_hi=hi)
        m = own == r
        s = _species(pgpu, g, x[:, m], xo[:, m], v[:, m], w[m], ids[m])
        s.set_current_density(1.0)
        g.current_zero(); g.current_add(s)
        grids.append(g); sps.append(s)
        if peer:
            hxs.append(halo.PeerHaloExchange(lay, r, g))
        else:
            hxs.append(halo.HaloExchange(lay, r, hub.view(r), halo.CapiGridBackend(g, torch.device("cuda", 0))))
    if peer:
        halo.PeerHaloExchange.connect_local(hxs)
        for rep in range(3):
            if rep:                                   # same deposit again: the exchange must give the same sums
                for g, s in zip(grids, sps):
                    g.current_zero(); g.current_add(s)
            for h in hxs:
                h.begin()
            for ph in range(hxs[0].nphase):
                for h in hxs:
                    h.send(ph)
                for h in hxs:
                    h.recv_add(ph)
    else:
        for ph in range(hxs[0].n_phases()):
            for h in hxs:
                h.begin(ph)
            for h in hxs:
                h.end(ph)
    worst = 0.0
    for r, g in enumerate(grids):
        g.current_finalize()          # folds the directions this box spans (none for 2x2 boxes)
        for c in range(3):
            lo, hi = g.field_bounds(c)
            a = g.current_get(c)
            (glo, _), ga = Jg[c]
            ii = np.mod(np.arange(lo[0], hi[0] + 1), NCELL[0]) - glo[0]
            jj = np.mod(np.arange(lo[1], hi[1] + 1), NCELL[1]) - glo[1]
            worst = max(worst, float(np.max(np.abs(a - ga[np.ix_(ii, jj)])) / np.max(np.abs(ga))))
    if peer:
        for h in hxs:
            h.destroy()
    for s in sps:
        s.destroy()
    for g in grids:
        g.destroy()
    assert worst < 1e-13, worst


@pytest.mark.parametrize("stag", [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("nbox", [(16, 16), (8, 16)])
def test_charge_density_add_exchange_matches_single_box(pgpu, nbox, stag):
    """setChargeDensity / OnFaces / OnNodes over several boxes (PicChargedSpecies.cpp:3083, :3120, :3141): deposit
    per box into the resident array, add-exchange of the ghost layers on the device, read -- against the same
    particles in one box spanning the (periodic) domain.  Twice, for both inbox slots."""
    lay = halo.BoxLayout(2, NCELL, nbox, NG, (1, 1))
    x, xo, v, w = _particles(2)
    ids = np.arange(w.size, dtype=np.uint64)
    g1 = pgpu.Grid(2, NCELL, XMIN, DX, NG, (1, 1))
    s1 = _species(pgpu, g1, x, xo, v, w, ids)
    ref, glo, _ = s1.charge_density(stag)
    s1.destroy(); g1.destroy()
    # owner by the old position (inside the domain); x is within half a cell of it, i.e. inside the owner's ghosts
    own = sum(np.floor((xo[d] - XMIN[d]) / (DX[d] * nbox[d])).astype(int) * (1 if d == 0 else lay.nb[0])
              for d in range(2))
    grids, sps, hxs = [], [], []
    for r in range(lay.world):
        lo, hi = lay.box(r)
        g = pgpu.Grid(2, NCELL, XMIN, DX, NG, (1, 1), box_lo=lo, box_hi=hi)
        m = own == r
        grids.append(g)
        sps.append(_species(pgpu, g, x[:, m], xo[:, m], v[:, m], w[m], ids[m]))
        hxs.append(halo.PeerHaloExchange(lay, r, g, rho_stag=stag))
    halo.PeerHaloExchange.connect_local(hxs)
    worst = 0.0
    for rep in range(2):
        for s in sps:
            s.charge_density_deposit(stag)
        for h in hxs:
            h.begin()
        for ph in range(hxs[0].nphase):
            for h in hxs:
                h.send(ph)
            for h in hxs:
                h.recv_add(ph)
        for g in grids:
            a, lo, hi = g.charge_density_get(stag)
            ii = np.mod(np.arange(lo[0], hi[0] + 1), NCELL[0]) - glo[0]
            jj = np.mod(np.arange(lo[1], hi[1] + 1), NCELL[1]) - glo[1]
            worst = max(worst, float(np.max(np.abs(a - ref[np.ix_(ii, jj)])) / np.max(np.abs(ref))))
    for h in hxs:
        h.destroy()
    for s in sps:
        s.destroy()
    for g in grids:
        g.destroy()
    assert worst < 1e-13, worst


def test_binomial_filter_of_J_and_rho(pgpu):
    """PicSpeciesInterface::filterJ / setChargeDensityOnNodes(use_filtering) (SpaceUtils.cpp:54-112): the device
    filter against the oracle's restatement bit for bit on one periodic box, and two boxes with the add-exchange
    in front of it against the one box."""
    from common import orc
    x, xo, v, w = _particles(5)
    ids = np.arange(w.size, dtype=np.uint64)
    g1 = pgpu.Grid(2, NCELL, XMIN, DX, NG, (1, 1))
    s1 = _species(pgpu, g1, x, xo, v, w, ids)
    s1.set_current_density(1.0)
    g1.current_zero(); g1.current_add(s1); g1.current_finalize()
    before = [g1.current_get(c) for c in range(3)]
    g1.current_filter(True, True)
    after = [g1.current_get(c) for c in range(3)]
    blo, bhi = (0, 0), (NCELL[0] - 1, NCELL[1] - 1)
    own = lambda a, st: a[NG:NG + NCELL[0] + st[0], NG:NG + NCELL[1] + st[1]]
    for c in range(3):
        lo, hi = g1.field_bounds(c)
        st = orc.E_STAG[2][c]
        f = orc.Fab(lo, hi, before[c].copy(order="F"))
        orc.binomial_filter(f, 2, blo, bhi, st)
        assert np.array_equal(own(after[c], st), own(f.a, st)), c
        # the periodic images follow their owners
        assert np.array_equal(after[c][NG - 1, NG:NG + NCELL[1]], after[c][NG + NCELL[0] - 1, NG:NG + NCELL[1]])
    # only the in-plane components
    g1.current_zero(); g1.current_add(s1); g1.current_finalize()
    g1.current_filter(True, False)
    assert np.array_equal(g1.current_get(2), before[2]) and np.array_equal(g1.current_get(0), after[0])
    # rho on nodes
    st = (1, 1)
    s1.charge_density_deposit(st)
    rho0, lo, hi = g1.charge_density_get(st)          # the same deposit (atomics: a second one may differ in the last bit)
    g1.charge_density_filter(st)
    rho1, _, _ = g1.charge_density_get(st)
    f = orc.Fab(lo, hi, rho0.copy(order="F"))
    orc.binomial_filter(f, 2, blo, bhi, st)
    assert np.array_equal(own(rho1, st), own(f.a, st))
    s1.destroy(); g1.destroy()
    # two boxes in x
    nbox = (16, 16)
    lay = halo.BoxLayout(2, NCELL, nbox, NG, (1, 1))
    ownr = np.floor((xo[0] - XMIN[0]) / (DX[0] * nbox[0])).astype(int)
    grids, sps, hxs = [], [], []
    for r in range(lay.world):
        lo, hi = lay.box(r)
        g = pgpu.Grid(2, NCELL, XMIN, DX, NG, (1, 1), box_lo=lo, box_hi=hi)
        m = ownr == r
        s = _species(pgpu, g, x[:, m], xo[:, m], v[:, m], w[m], ids[m])
        s.set_current_density(1.0)
        g.current_zero(); g.current_add(s)
        grids.append(g); sps.append(s); hxs.append(halo.PeerHaloExchange(lay, r, g))
    halo.PeerHaloExchange.connect_local(hxs)
    for h in hxs:
        h.begin()
    for ph in range(hxs[0].nphase):
        for h in hxs:
            h.send(ph)
        for h in hxs:
            h.recv_add(ph)
    worst = 0.0
    for r, g in enumerate(grids):
        g.current_finalize()
        g.current_filter(True, True)
        blo_r, bhi_r = lay.box(r)
        for c in range(3):
            lo, hi = g.field_bounds(c)
            st = orc.E_STAG[2][c]
            a = g.current_get(c)
            sl = (slice(blo_r[0] - lo[0], bhi_r[0] + st[0] - lo[0] + 1), slice(blo_r[1] - lo[1], bhi_r[1] + st[1] - lo[1] + 1))
            ii = np.mod(np.arange(blo_r[0], bhi_r[0] + st[0] + 1), NCELL[0]) + NG
            jj = np.mod(np.arange(blo_r[1], bhi_r[1] + st[1] + 1), NCELL[1]) + NG
            worst = max(worst, float(np.max(np.abs(a[sl] - after[c][np.ix_(ii, jj)])) / np.max(np.abs(after[c]))))
    for h in hxs:
        h.destroy()
    for s in sps:
        s.destroy()
    for g in grids:
        g.destroy()
    assert worst < 1e-13, worst


@pytest.mark.parametrize("route", ["mailbox", "peer"])
def test_migration_on_device(pgpu, route):
    """route=mailbox: mark / pack / append around a message layer (the NCCL route's kernels);
    route=peer: pgpu_migrator_* -- leavers stored straight into the neighbours' inboxes, counts on the device."""
    import torch
    nbox = (16, 8)
    lay = halo.BoxLayout(2, NCELL, nbox, NG, (1, 1))
    x, xo, v, w = _particles(9, n=30000)
    L = np.array([nc * h for nc, h in zip(NCELL, DX)])
    ids = np.arange(w.size, dtype=np.uint64) + 1000
    box_of = lambda p: sum(np.floor((p[d] - XMIN[d]) / (DX[d] * nbox[d])).astype(int) * (1 if d == 0 else lay.nb[0])
                           for d in range(2))
    own_old = box_of(xo)
    hub = halo.LoopComm(lay.world)
    grids, sps, migs = [], [], []
    dev = torch.device("cuda", 0)
    for r in range(lay.world):
        lo, hi = lay.box(r)
        g = pgpu.Grid(2, NCELL, XMIN, DX, NG, (1, 1), box_lo=lo, box_hi=hi)
        m = own_old == r
        s = _species(pgpu, g, x[:, m], xo[:, m], v[:, m], w[m], ids[m])
        s.apply_bcs((1, 1), (1, 1))                       # periodic wrap of x and xold
        grids.append(g); sps.append(s)
        if route == "peer":
            migs.append(halo.PeerMigration(lay, r, s, capacity=4096))
        else:
            migs.append(halo.Migration(lay, r, hub.view(r), halo.CapiSpeciesBackend(s, dev)))
    if route == "peer":
        halo.PeerMigration.connect_local(migs)
        n_before = sum(s.n for s in sps)
        moved = halo.migrate_all_peer(migs)
        assert moved > 200 and all(m.lost == 0 for m in migs) and sum(s.n for s in sps) == n_before
    else:
        counts = [m.begin_counts().numpy() for m in migs]
        moved = sum(int(c.sum()) for c in counts)
        assert moved > 200 and all(m.lost == 0 for m in migs)
        for m in migs:
            m.begin_payload(counts)
        assert sum(m.end() for m in migs) == moved
    xw = np.array(XMIN)[:, None] + np.mod(x - np.array(XMIN)[:, None], L[:, None])
    own_new = box_of(xw)
    order = np.argsort(ids)
    total = 0
    for r, s in enumerate(sps):
        got = s.download()
        total += got["w"].size
        assert np.array_equal(np.sort(got["id"]), np.sort(ids[own_new == r]))
        src = order[np.searchsorted(ids[order], got["id"])]
        # the records arrive unchanged; positions carry the periodic wrap of applyBCs
        assert np.array_equal(got["w"], w[src]) and np.array_equal(got["v"], v[:, src])
        assert np.max(np.abs(got["x"] - xw[:, src])) < 1e-12
        lo, hi = lay.box(r)
        for d in range(2):
            c = np.floor((got["x"][d] - XMIN[d]) / DX[d])
            assert c.min() >= lo[d] and c.max() <= hi[d]
        # nobody leaves any more
        if route != "peer":
            assert int(migs[r].begin_counts().sum()) == 0
    assert total == w.size
    if route == "peer":
        assert halo.migrate_all_peer(migs) == 0          # second round (other parity): nothing moves
        assert sum(s.n for s in sps) == w.size
        # overflow is an error, not silent loss: capacity 4 cannot take the leavers of a fresh start
        for m in migs:
            m.destroy()
    for s in sps:
        s.destroy()
    for g in grids:
        g.destroy()


# ------------------------------------------------------------------------------------------------
# the CUDA IPC route: two PROCESSES (both on cuda:0 -- IPC does not care), handles over gloo
# ------------------------------------------------------------------------------------------------
def _ipc_worker(rank, world, port, q):
    import os
    import torch
    import torch.distributed as dist
    from picnic_b200 import capi
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        capi.load(); capi.init(0)
        nbox = (16, 16)
        lay = halo.BoxLayout(2, NCELL, nbox, NG, (1, 1))          # 2 x 1 boxes; y is folded locally
        x, xo, v, w = _particles(2)
        ids = np.arange(w.size, dtype=np.uint64)
        own = np.floor((xo[0] - XMIN[0]) / (DX[0] * nbox[0])).astype(int)
        lo, hi = lay.box(rank)
        g = capi.Grid(2, NCELL, XMIN, DX, NG, (1, 1), box_lo=lo, box_hi=hi)
        m = own == rank
        s = _species(capi, g, x[:, m], xo[:, m], v[:, m], w[m], ids[m])
        s.set_current_density(1.0)
        hx = halo.PeerHaloExchange(lay, rank, g)
        hx.connect_ipc(halo.DistComm(rank, world))
        out = None
        for rep in range(3):                                      # both parity slots, growing sequence numbers
            g.current_zero(); g.current_add(s)
            hx.add_exchange()
            g.current_finalize()
            out = [g.current_get(c) for c in range(3)]
            dist.barrier()
        bounds = [g.field_bounds(c) for c in range(3)]
        dist.barrier()                                            # nobody tears its inbox down while a peer may write
        hx.destroy(); s.destroy(); g.destroy(); capi.finalize()
        q.put((rank, bounds, out))
    finally:
        dist.destroy_process_group()


def test_peer_exchange_two_processes_ipc(pgpu):
    import socket
    import torch.multiprocessing as mp
    x, xo, v, w = _particles(2)
    ids = np.arange(w.size, dtype=np.uint64)
    g1 = pgpu.Grid(2, NCELL, XMIN, DX, NG, (1, 1))
    s1 = _species(pgpu, g1, x, xo, v, w, ids)
    s1.set_current_density(1.0)
    g1.current_zero(); g1.current_add(s1); g1.current_finalize()
    Jg = [(g1.field_bounds(c), g1.current_get(c)) for c in range(3)]
    s1.destroy(); g1.destroy()
    sk = socket.socket(); sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]; sk.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ipc_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for (rank, bounds, out) in res:
        for c in range(3):
            (lo, hi), a = bounds[c], out[c]
            (glo, _), ga = Jg[c]
            ii = np.mod(np.arange(lo[0], hi[0] + 1), NCELL[0]) - glo[0]
            jj = np.mod(np.arange(lo[1], hi[1] + 1), NCELL[1]) - glo[1]
            assert float(np.max(np.abs(a - ga[np.ix_(ii, jj)])) / np.max(np.abs(ga))) < 1e-13, (rank, c)
