"""TEST / BENCH INFRASTRUCTURE -- never imported by the product path (picnic_b200/).

The CPU arm of bench.py (`--impl reference` and the `cpu_baseline` leg): the oracle (oracle/*.cpp, the
operation-order restatement of the reference's particle routines) run the way the reference runs, as R
box-owning workers.  Every worker owns one square box of a periodic domain (System.cpp:169-245), its particles
(SoA) and its ghosted E, B and J arrays, and does per step what PicSpeciesInterface::preRHSOp
(PicSpeciesInterface.cpp:899-994) does per nonlinear evaluation -- particle-Picard advance
(PicChargedSpecies.cpp:1614-1716), setCurrentDensity, ghost ADD-exchange of J with the neighbouring boxes
(LevelData::exchange + addExchange stand-in: shared-memory mailbox) -- and once per step the second half
advance, periodic applyBCs and the remap of leavers to the owning box (ParticleData::remapOutcast stand-in).
The workers are threads; the oracle's C routines release the GIL, the numpy glue around them is a few per
cent of a step.

The index arithmetic of the exchanges is picnic_b200/halo.py's (HaloExchange / Migration are backend-neutral);
the numpy backends below implement its pack / unpack-add / mark / pack-leavers / append contract.
"""
import collections
import math
import threading
import time

import numpy as np
import torch

from picnic_b200 import decks, halo
from . import oracle as orc


# ------------------------------------------------------------------------------------------------
# numpy backends of picnic_b200.halo (also used by tests/test_halo_cpu.py)
# ------------------------------------------------------------------------------------------------
class NumpyGridBackend:
    def __init__(self, layout, rank):
        self.device = "cpu"
        self.arr = []
        for stag in halo.STAG_J[layout.D]:
            lo, hi = layout.array_bounds(rank, stag)
            shape = tuple(h - l + 1 for l, h in zip(lo, hi))
            self.arr.append((lo, hi, np.zeros(shape, order="F")))

    def new_buffer(self, count):
        return torch.empty(count, dtype=torch.float64)

    def _view(self, comp, lo, hi):
        alo, _, a = self.arr[comp]
        return a[tuple(slice(l - al, h - al + 1) for l, h, al in zip(lo, hi, alo))]

    def pack(self, comp, lo, hi, buf):
        buf.copy_(torch.from_numpy(np.ascontiguousarray(self._view(comp, lo, hi).ravel(order="F"))))

    def unpack_add(self, comp, lo, hi, buf):
        v = self._view(comp, lo, hi)
        v += buf.numpy().reshape(v.shape, order="F")

    def sync(self):
        pass


class NumpySpeciesBackend:
    """Particles of one box; ownership rule of pgpu_species_mark_leavers."""

    def __init__(self, layout, rank, x, xold, v, vold, w, ids, xmin, dx):
        self.layout, self.rank, self.device = layout, rank, "cpu"
        self.D = layout.D
        self.nw = 2 * self.D + 8
        self.p = dict(x=x.copy(), xold=xold.copy(), v=v.copy(), vold=vold.copy(), w=w.copy(), id=ids.copy())
        self.xmin, self.dx = np.asarray(xmin, float), np.asarray(dx, float)
        self.codes = None

    @property
    def n(self):
        return self.p["w"].size

    def owner_codes(self):
        lay = self.layout
        my = lay.coords(self.rank)
        code = np.zeros(self.n, dtype=np.int64)
        mul = 1
        lost = np.zeros(self.n, dtype=bool)
        for d in range(self.D):
            b = np.floor((self.p["x"][d] - self.xmin[d]) / (self.dx[d] * lay.nbox[d])).astype(np.int64)
            diff = b - my[d]
            if lay.periodic[d]:
                diff = np.where(diff > 1, diff - lay.nb[d], diff)
                diff = np.where(diff < -1, diff + lay.nb[d], diff)
            lost |= (b < 0) | (b >= lay.nb[d]) | (np.abs(diff) > 1)
            code += (diff + 1) * mul
            mul *= 3
        if self.D == 1:
            code += 3
        return np.where(lost, 9, code)

    def mark_leavers(self):
        self.codes = self.owner_codes()
        return np.bincount(np.where(self.codes == 4, 10, self.codes), minlength=11)[:10].astype(np.int64)

    def new_buffer(self, nrec):
        return torch.empty(max(nrec, 1) * self.nw, dtype=torch.float64)

    def _records(self, idx):
        p = self.p
        cols = [p["x"][d][idx] for d in range(self.D)] + [p["xold"][d][idx] for d in range(self.D)]
        cols += [p["v"][c][idx] for c in range(3)] + [p["vold"][c][idx] for c in range(3)]
        cols += [p["w"][idx], p["id"][idx].view(np.float64)]
        return np.stack(cols, axis=1)

    def pack_leavers(self, buf):
        order = np.argsort(self.codes, kind="stable")
        order = order[(self.codes[order] != 4) & (self.codes[order] < 9)]
        rec = self._records(order)
        buf[:rec.size].copy_(torch.from_numpy(np.ascontiguousarray(rec).ravel()))
        keep = self.codes == 4
        if not keep.all():
            for k, a in self.p.items():
                self.p[k] = a[..., keep].copy()

    def append(self, nrec, buf):
        rec = buf[:nrec * self.nw].numpy().reshape(nrec, self.nw)
        D, p = self.D, self.p
        p["x"] = np.concatenate([p["x"], rec[:, 0:D].T], axis=1)
        p["xold"] = np.concatenate([p["xold"], rec[:, D:2 * D].T], axis=1)
        p["v"] = np.concatenate([p["v"], rec[:, 2 * D:2 * D + 3].T], axis=1)
        p["vold"] = np.concatenate([p["vold"], rec[:, 2 * D + 3:2 * D + 6].T], axis=1)
        p["w"] = np.concatenate([p["w"], rec[:, 2 * D + 6]])
        p["id"] = np.concatenate([p["id"], np.ascontiguousarray(rec[:, 2 * D + 7]).view(np.uint64)])

    def sync(self):
        pass


# ------------------------------------------------------------------------------------------------
# shared-memory communicator between worker threads (the reference's MPI ranks)
# ------------------------------------------------------------------------------------------------
class ThreadHub:
    """Mailbox keyed by (source, destination, tag), FIFO per key, plus a barrier-based all-gather."""

    def __init__(self, world):
        self.world = world
        self.box = collections.defaultdict(collections.deque)
        self.cv = threading.Condition()
        self.gathered = [None] * world
        self.barrier = threading.Barrier(world)

    def view(self, rank):
        return _ThreadView(self, rank)


class _ThreadView:
    def __init__(self, hub, rank):
        self.hub, self.rank, self.world = hub, rank, hub.world
        self.recvs = []

    def post(self, sends, recvs):
        with self.hub.cv:
            for (peer, tag, t) in sends:
                self.hub.box[(self.rank, peer, tag)].append(t.clone())
            self.hub.cv.notify_all()
        self.recvs = recvs

    def wait(self):
        for (peer, tag, t) in self.recvs:
            key = (peer, self.rank, tag)
            with self.hub.cv:
                self.hub.cv.wait_for(lambda: len(self.hub.box[key]) > 0)
                src = self.hub.box[key].popleft()
            t.copy_(src)
        self.recvs = []

    def all_gather(self, t):
        self.hub.gathered[self.rank] = t.clone()
        self.hub.barrier.wait()
        out = list(self.hub.gathered)
        self.hub.barrier.wait()
        return out


# ------------------------------------------------------------------------------------------------
# one box-owning worker
# ------------------------------------------------------------------------------------------------
class BoxWorker:
    def __init__(self, deck, layout, rank, hub, E0, B0, n_outer, eps_outer):
        self.deck, self.lay, self.rank = deck, layout, rank
        self.lo, self.hi = layout.box(rank)
        self.geom = orc.make_geom(2, deck.xmin, deck.xmax, deck.dx, deck.nghost)
        E, B = decks.analytic_fields(deck, self.lo, self.hi, E0=E0, B0=B0)
        # one field set per nonlinear evaluation of a step, as the GPU arm uploads them
        self.fields = []
        for j in range(n_outer):
            s = 1.0 + eps_outer[j % len(eps_outer)]
            self.fields.append(([orc.Fab(l, h, a * s) for (l, h, a) in E], [orc.Fab(l, h, a * s) for (l, h, a) in B]))
        self.grid = NumpyGridBackend(layout, rank)
        self.J = [orc.Fab(l, h, a) for (l, h, a) in self.grid.arr]
        for f, (_, _, a) in zip(self.J, self.grid.arr):
            assert f.a is a or np.shares_memory(f.a, a)
        comm = hub.view(rank)
        self.halo = halo.HaloExchange(layout, rank, comm, self.grid)
        self.fold = [d for d in range(2) if layout.nb[d] == 1]
        rng = np.random.default_rng(deck.seed + 1000 * rank)
        self.species, self.migration = [], []
        for sdef in deck.species:
            p = decks.load_species(deck, sdef, self.lo, self.hi, rng)
            be = NumpySpeciesBackend(layout, rank, p["x"], p["x"], p["v"], p["v"], p["w"],
                                     np.asarray(p["id"], dtype=np.uint64), deck.xmin, deck.dx)
            self.species.append((sdef, be))
            self.migration.append(halo.Migration(layout, rank, hub.view(rank), be))
        self.L = [n * h for n, h in zip(deck.ncell, deck.dx)]
        self.apply_its = 0
        self.unconverged = 0
        self.migrated = 0

    def step(self, n_outer):
        deck, lib = self.deck, orc.lib()
        units = 0
        for _, be in self.species:
            be.p["xold"][...] = be.p["x"]
            be.p["vold"][...] = be.p["v"]
        for j in range(n_outer):
            Ef, Bf = self.fields[j]
            for f in self.J:
                f.a[...] = 0.0
            for sdef, be in self.species:
                p = be.p
                rc, its, unc, _ = orc.advance_particles_iteratively(
                    self.geom, deck.interp_E, p["x"], p["xold"], p["v"], p["vold"], Ef, Bf,
                    sdef.fnorm_const(deck.units), deck.cnorm_dt, deck.rtol, deck.iter_max)
                assert rc == 0
                self.apply_its += its
                self.unconverged += unc
                orc.deposit_current(self.geom, deck.interp_J, p["x"], p["xold"], p["v"], p["w"], deck.cnorm_dt, self.J)
                units += be.n
            self.halo.add_exchange()
            for comp, stag in enumerate(halo.STAG_J[2]):     # a direction one box wide folds onto itself
                if self.fold:
                    per = [1 if d in self.fold else 0 for d in range(2)]
                    orc.fold_periodic(self.J[comp], 2, stag, self.lo, self.hi, per)
        for k, (sdef, be) in enumerate(self.species):
            p = be.p
            lib.orc_advance_velocities_2nd_half(be.n, orc._ptr(p["v"]), orc._ptr(p["vold"]))
            lib.orc_advance_positions_2nd_half(2, be.n, orc._ptr(p["x"]), orc._ptr(p["xold"]))
            for d in range(2):                                # periodic applyBCs
                xd = p["x"][d]
                shift = np.where(xd < deck.xmin[d], self.L[d], 0.0) - np.where(xd >= deck.xmin[d] + self.L[d], self.L[d], 0.0)
                if shift.any():
                    xd += shift
                    p["xold"][d] += shift
            self.migrated += self.migration[k].migrate()      # collective: every worker, species by species
        return units


def worker_grid(nthreads):
    """px x py boxes for at most nthreads workers, as square as the count allows."""
    py = max(1, int(math.isqrt(max(nthreads, 1))))
    px = max(1, nthreads // py)
    return px, py


def probe_rate(deck_fn, E0B0_fn, n_outer, eps_outer, bn=24):
    """particle-advances/s of ONE worker on a small box (sizes the sample)."""
    deck = deck_fn((bn, bn))
    lay = halo.BoxLayout(2, deck.ncell, (bn, bn), deck.nghost, (1, 1))
    E0, B0 = E0B0_fn(deck)
    w = BoxWorker(deck, lay, 0, ThreadHub(1), E0, B0, n_outer, eps_outer)
    t0 = time.perf_counter()
    units = w.step(n_outer)
    return units / (time.perf_counter() - t0)


def run(deck_fn, E0B0_fn, nthreads, steps, warmup, n_outer, eps_outer, bn):
    """`steps` timed steps (after `warmup`) of px x py boxes of bn^2 cells, one thread per box.
    deck_fn((n0, n1)) returns the deck of a periodic domain of n0 x n1 cells."""
    px, py = worker_grid(nthreads)
    R = px * py
    deck = deck_fn((bn * px, bn * py))
    lay = halo.BoxLayout(2, deck.ncell, (bn, bn), deck.nghost, (1, 1))
    assert lay.world == R
    E0, B0 = E0B0_fn(deck)
    hub = ThreadHub(R)
    workers = [BoxWorker(deck, lay, r, hub, E0, B0, n_outer, eps_outer) for r in range(R)]
    n0 = sum(be.n for w in workers for _, be in w.species)
    units = [0] * R
    tmark = [0.0, 0.0]
    errors = []

    def body(r):
        try:
            w = workers[r]
            for s in range(warmup + steps):
                if s == warmup:
                    hub.barrier.wait()
                    if r == 0:
                        tmark[0] = time.perf_counter()
                        for ww in workers:
                            ww.apply_its = 0
                    hub.barrier.wait()
                u = w.step(n_outer)
                if s >= warmup:
                    units[r] += u
            hub.barrier.wait()
            if r == 0:
                tmark[1] = time.perf_counter()
        except BaseException as e:      # a dead worker would leave the others in a barrier
            errors.append(repr(e))
            hub.barrier.abort()

    threads = [threading.Thread(target=body, args=(r,)) for r in range(R)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise RuntimeError("CPU worker failed: " + errors[0])
    n1 = sum(be.n for w in workers for _, be in w.species)
    assert n1 == n0, "migration lost particles: %d -> %d" % (n0, n1)
    sec = tmark[1] - tmark[0]
    tot = sum(units)
    return {"units": tot, "seconds": sec, "workers": R, "boxes": "%dx%d" % (px, py), "box_cells": bn,
            "particles": n0, "migrated": sum(w.migrated for w in workers),
            "mean_picard_passes": sum(w.apply_its for w in workers) / max(tot, 1)}
