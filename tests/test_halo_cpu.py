"""Host-side logic of the multi-box exchanges (picnic_b200/halo.py) on CPU: index boxes, the
direction-by-direction add-exchange and the migration, first through the in-process mailbox
(4 boxes), then with two real processes over gloo.  The device side is replaced by the numpy
stand-ins of tests/halo_numpy.py; the CUDA kernels are covered by tests/test_gpu_halo.py."""
import os
import socket

import numpy as np
import pytest
import torch

from common import orc, decks  # noqa: F401
from picnic_b200 import halo
from halo_numpy import NumpyGridBackend, NumpySpeciesBackend

NCELL, NBOX, NG = (16, 16), (8, 8), 2
DX, XMIN = (0.25, 0.5), (0.0, -1.0)


def _particles(seed, n=4000):
    rng = np.random.default_rng(seed)
    L = np.array([nc * h for nc, h in zip(NCELL, DX)])
    xo = np.array(XMIN)[:, None] + rng.random((2, n)) * L[:, None]
    x = xo + (rng.random((2, n)) - 0.5) * np.array(DX)[:, None] * 0.8
    v = rng.standard_normal((3, n)) * 0.05
    w = rng.random(n) + 0.5
    return x, xo, v, w


def _deposit_box(lo, hi, x, xo, v, w):
    """CC1 deposit of the given particles into the ghosted arrays of box lo..hi (oracle)."""
    xmax = tuple(x0 + nc * h for x0, nc, h in zip(XMIN, NCELL, DX))
    geom = orc.make_geom(2, XMIN, xmax, DX, NG)
    J = [orc.fab_for(lo, hi, NG, s) for s in orc.E_STAG[2]]
    if w.size:
        rc = orc.deposit_current(geom, orc.CC1, np.ascontiguousarray(x), np.ascontiguousarray(xo),
                                 np.ascontiguousarray(v), np.ascontiguousarray(w), 1.0, J)
        assert rc == 0
    return J


def _owner(xo, layout):
    """Box of the start-of-step position (the reference's per-box particle membership)."""
    b = [np.floor((xo[d] - XMIN[d]) / (DX[d] * NBOX[d])).astype(int) for d in range(2)]
    return b[0] + b[1] * layout.nb[0]


def _global_reference(x, xo, v, w):
    """The same particles deposited on ONE box spanning the domain, ghosts folded periodically."""
    J = _deposit_box((0, 0), (NCELL[0] - 1, NCELL[1] - 1), x, xo, v, w)
    for c, f in enumerate(J):
        orc.fold_periodic(f, 2, orc.E_STAG[2][c], (0, 0), (NCELL[0] - 1, NCELL[1] - 1), (1, 1))
    return J


def _check_against_global(layout, rank, be, Jg):
    """Every entry of this box's arrays (ghosts included) equals the folded global value at the
    periodic image of its index."""
    worst = 0.0
    for c, stag in enumerate(halo.STAG_J[2]):
        lo, hi, a = be.arr[c]
        g = Jg[c]
        ii = np.mod(np.arange(lo[0], hi[0] + 1), NCELL[0]) - g.lo[0]
        jj = np.mod(np.arange(lo[1], hi[1] + 1), NCELL[1]) - g.lo[1]
        want = g.a[np.ix_(ii, jj)]
        worst = max(worst, float(np.max(np.abs(a - want)) / np.max(np.abs(g.a))))
    return worst


def test_layout_boxes_and_overlaps():
    lay = halo.BoxLayout(2, NCELL, NBOX, NG, (1, 1))
    assert lay.world == 4 and lay.box(3) == ((8, 8), (15, 15))
    assert lay.neighbor(0, 0, -1) == 1 and lay.neighbor(0, 1, +1) == 2      # periodic wrap, 2 boxes per dir
    assert lay.neighbor_code(0, 8) == 3 and lay.neighbor_code(3, 0) == 0
    lo, hi = lay.overlap(0, (1, 0), 0, +1)           # nodal in x: 2*2+1 layers around the face at node 8
    assert (lo[0], hi[0]) == (6, 10) and (lo[1], hi[1]) == (-2, 9)
    wall = halo.BoxLayout(2, NCELL, NBOX, NG, (0, 1))
    assert wall.neighbor(0, 0, -1) is None and wall.neighbor(1, 0, +1) is None


def test_add_exchange_four_boxes_mailbox():
    lay = halo.BoxLayout(2, NCELL, NBOX, NG, (1, 1))
    x, xo, v, w = _particles(3)
    own = _owner(xo, lay)
    hub = halo.LoopComm(lay.world)
    bes, hxs = [], []
    for r in range(lay.world):
        be = NumpyGridBackend(lay, r)
        m = own == r
        J = _deposit_box(*lay.box(r), x[:, m], xo[:, m], v[:, m], w[m])
        for c in range(3):
            be.arr[c][2][...] = J[c].a
        bes.append(be)
        hxs.append(halo.HaloExchange(lay, r, hub.view(r), be))
    for ph in range(hxs[0].n_phases()):
        for h in hxs:
            h.begin(ph)
        for h in hxs:
            h.end(ph)
    Jg = _global_reference(x, xo, v, w)
    for r in range(lay.world):
        assert _check_against_global(lay, r, bes[r], Jg) < 1e-13


def test_migration_four_boxes_mailbox():
    lay = halo.BoxLayout(2, NCELL, NBOX, NG, (1, 1))
    x, xo, v, w = _particles(5, n=3000)
    L = np.array([nc * h for nc, h in zip(NCELL, DX)])
    xw = np.array(XMIN)[:, None] + np.mod(x - np.array(XMIN)[:, None], L[:, None])     # periodic applyBCs
    ids = np.arange(w.size, dtype=np.uint64) + 7
    own_old = _owner(xo, lay)
    hub = halo.LoopComm(lay.world)
    migs, bes = [], []
    for r in range(lay.world):
        m = own_old == r
        be = NumpySpeciesBackend(lay, r, xw[:, m], xo[:, m], v[:, m], v[:, m] * 0.5, w[m], ids[m], XMIN, DX)
        bes.append(be)
        migs.append(halo.Migration(lay, r, hub.view(r), be))
    counts = [m.begin_counts().numpy() for m in migs]
    assert sum(int(c.sum()) for c in counts) > 50                # the test moves particles
    for m in migs:
        m.begin_payload(counts)
    n_in = sum(m.end() for m in migs)
    assert n_in == sum(int(c.sum()) for c in counts)
    assert sum(be.n for be in bes) == w.size
    own_new = _owner(xw, lay)
    for r, be in enumerate(bes):
        assert np.all(be.owner_codes() == 4)                     # everybody is at home now
        want = np.sort(ids[own_new == r])
        assert np.array_equal(np.sort(be.p["id"]), want)
        k = np.argsort(be.p["id"])
        src = np.argsort(ids)[np.searchsorted(np.sort(ids), be.p["id"][k])]
        assert np.array_equal(be.p["x"][:, k], xw[:, src]) and np.array_equal(be.p["w"][k], w[src])


# ------------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lay = halo.BoxLayout(2, NCELL, (8, 16), NG, (1, 1))         # 2 x 1 boxes: y is folded locally
        x, xo, v, w = _particles(11)
        b0 = np.floor((xo[0] - XMIN[0]) / (DX[0] * 8)).astype(int)
        m = b0 == rank
        be = NumpyGridBackend(lay, rank)
        J = _deposit_box(*lay.box(rank), x[:, m], xo[:, m], v[:, m], w[m])
        for c in range(3):
            be.arr[c][2][...] = J[c].a
        hx = halo.HaloExchange(lay, rank, halo.DistComm(rank, world), be)
        hx.add_exchange()
        # the direction the box spans is folded locally (what pgpu_current_finalize does)
        lo, hi = lay.box(rank)
        for c in range(3):
            f = orc.Fab(be.arr[c][0], be.arr[c][1], be.arr[c][2])
            orc.fold_periodic(f, 2, orc.E_STAG[2][c], lo, hi, (0, 1))
            be.arr[c][2][...] = f.a
        Jg = _global_reference(x, xo, v, w)
        err = _check_against_global(lay, rank, be, Jg)
        # migration over gloo
        L = np.array([nc * h for nc, h in zip(NCELL, DX)])
        xw = np.array(XMIN)[:, None] + np.mod(x - np.array(XMIN)[:, None], L[:, None])
        ids = np.arange(w.size, dtype=np.uint64)
        sb = NumpySpeciesBackend(lay, rank, xw[:, m], xo[:, m], v[:, m], v[:, m], w[m], ids[m], XMIN, DX)
        mg = halo.Migration(lay, rank, halo.DistComm(rank, world), sb)
        n_in = mg.migrate()
        home = bool(np.all(sb.owner_codes() == 4))
        tot = torch.tensor([sb.n], dtype=torch.int64)
        dist.all_reduce(tot)
        q.put((rank, err, n_in, home, int(tot.item()), int(w.size)))
    finally:
        dist.destroy_process_group()


def test_add_exchange_and_migration_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for (rank, err, n_in, home, tot, n) in res:
        assert err < 1e-13, (rank, err)
        assert n_in > 0 and home and tot == n
