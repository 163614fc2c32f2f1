"""Shared builders for the parity tests: a small periodic problem whose inputs go to
both the CPU oracle (oracle/) and the CUDA library (picnic_b200/) unchanged."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402
from picnic_b200 import decks  # noqa: E402

INTERPS = {"CIC": 0, "TSC": 1, "CC0": 2, "CC1": 3}


class Problem:
    """Periodic box of ncell cells with nghost ghost layers, random smooth fields and
    particles whose half-step displacement reaches `max_disp` cells."""

    def __init__(self, D, ncell, dx, xmin, nghost, n, seed=0, max_disp=1.2, E0=1.0, B0=1.0):
        rng = np.random.default_rng(seed)
        self.D, self.ncell, self.dx, self.xmin, self.nghost, self.n = D, tuple(ncell), tuple(dx), tuple(xmin), nghost, n
        self.xmax = tuple(x0 + nc * h for x0, nc, h in zip(xmin, ncell, dx))
        self.box_lo = (0,) * D
        self.box_hi = tuple(nc - 1 for nc in ncell)
        self.geom = orc.make_geom(D, self.xmin, self.xmax, self.dx, nghost)
        # fields: random values, made periodic so ghosts are consistent images
        self.E = [self._random_field(rng, st, E0) for st in orc.E_STAG[D]]
        self.B = [self._random_field(rng, st, B0) for st in orc.B_STAG[D]]
        # particles
        L = np.array([nc * h for nc, h in zip(ncell, dx)])
        x0 = np.array(xmin)[:, None] + rng.random((D, n)) * L[:, None]
        disp = (rng.random((D, n)) * 2 - 1) * max_disp * np.array(dx)[:, None] * 0.5
        self.xold = np.ascontiguousarray(x0)
        self.x = np.ascontiguousarray(x0 + disp)          # xbar
        self.vold = np.ascontiguousarray(rng.standard_normal((3, n)) * 0.05)
        self.v = np.ascontiguousarray(self.vold + rng.standard_normal((3, n)) * 0.01)
        self.w = np.ascontiguousarray(rng.random(n) + 0.5)

    def _random_field(self, rng, stag, amp):
        D = self.D
        f = orc.fab_for(self.box_lo, self.box_hi, self.nghost, stag)
        shape = f.a.shape
        core = rng.standard_normal(self.ncell) * amp
        # periodic extension: value at index i is core[(i) mod ncell]
        idx = [np.mod(np.arange(f.lo[d], f.hi[d] + 1), self.ncell[d]) for d in range(D)]
        if D == 1:
            f.a[:] = core[idx[0]]
        else:
            f.a[:, :] = core[np.ix_(idx[0], idx[1])]
        assert f.a.shape == shape
        return f

    def fields_for_gpu(self):
        E = [(f.lo, f.hi, f.a) for f in self.E]
        B = [(f.lo, f.hi, f.a) for f in self.B]
        return E, B

    def new_J(self):
        return [orc.fab_for(self.box_lo, self.box_hi, self.nghost, st) for st in orc.E_STAG[self.D]]


def make_gpu(pgpu, prob, interp, rtol=1e-12, iter_max=21, order_swap=0, fnorm=1.0, cvac_norm=1.0,
             charge=-1.0, mass=1.0, volume_scale=1.0, interp_N=1, periodic=None):
    periodic = [1] * prob.D if periodic is None else periodic
    grid = pgpu.Grid(prob.D, prob.ncell, prob.xmin, prob.dx, prob.nghost, periodic, volume_scale=volume_scale)
    E, B = prob.fields_for_gpu()
    grid.set_fields(E, B)
    sp = pgpu.Species(grid, mass, charge, fnorm, cvac_norm, interp_N=interp_N, interp_J=interp, interp_E=interp,
                      rtol=rtol, iter_max=iter_max, order_swap=order_swap)
    sp.upload(prob.x, prob.v, prob.w, xold=prob.xold, vold=prob.vold, ids=np.arange(prob.n, dtype=np.uint64))
    return grid, sp


def rel_err(a, b, scale=None):
    a, b = np.asarray(a), np.asarray(b)
    s = np.max(np.abs(b)) if scale is None else scale
    return float(np.max(np.abs(a - b)) / (s if s > 0 else 1.0))
