#!/usr/bin/env python
"""Device timing of setMassMatrices on the C4 stand-in (1D, 250 000 cells x 200 ppc x 2 species = 1e8 particles, CC1):
the generic one-thread-per-particle kernel with warp-aggregated reductions (there is no 1D run kernel)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from picnic_b200 import capi, decks

capi.init(0)
lib = capi.load()
deck = decks.deck_c4()
deck.dt = 0.1
lo, hi = (0,), (deck.ncell[0] - 1,)
E, B = decks.analytic_fields(deck, lo, hi, E0=3.0e7, B0=5.0e8)
grid = capi.Grid(1, deck.ncell, deck.xmin, deck.dx, deck.nghost, (1,), volume_scale=deck.volume_scale)
grid.set_fields(E, B)
rng = np.random.default_rng(3)
sps = []
for sdef in deck.species:
    p = decks.load_species(deck, sdef, lo, hi, rng)
    sp = capi.Species(grid, sdef.mass, sdef.charge, sdef.fnorm_const(deck.units), deck.units.cvac_norm,
                      interp_N=deck.interp_N, interp_J=deck.interp_J, interp_E=deck.interp_E, rtol=deck.rtol,
                      iter_max=deck.iter_max)
    sp.upload(p["x"], p["v"], p["w"], ids=p["id"])
    sp.bin_particles()
    sps.append(sp)
    del p
n = sum(sp.n for sp in sps)
for sp in sps:
    capi.check(lib.pgpu_advance_particles_iteratively(sp.h, deck.dt, 1, None))
nc = grid.mass_matrices_init(3)


def set_mm():
    grid.mass_matrices_zero()
    for sp in sps:
        sp.accumulate_mass_matrices(deck.dt)
    grid.mass_matrices_save_E0()


set_mm(); capi.check(lib.pgpu_synchronize())
t0 = time.perf_counter()
for _ in range(3):
    set_mm()
capi.check(lib.pgpu_synchronize())
t = (time.perf_counter() - t0) / 3
capi.check(lib.pgpu_picard_totals(None, None, None, 1))
print("C4 1D setMassMatrices: %d particles, %d sigma components, %.2f ms = %.3e particles/s" % (
    n, int(sum(int(a) * int(b) for a, b in nc)), t * 1e3, n / t))
t0 = time.perf_counter()
for _ in range(3):
    grid.compute_J_from_mass_matrices()
capi.check(lib.pgpu_synchronize())
print("computeJfromMassMatrices: %.3f ms" % ((time.perf_counter() - t0) / 3 * 1e3))
capi.finalize()
