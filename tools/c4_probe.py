#!/usr/bin/env python
"""C4 (1D, 1e8 particles) fused advance+deposit vs advance alone vs deposit alone: where the generic kernel's time goes."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from picnic_b200 import capi, decks
capi.init(0)
lib = capi.load()
deck = decks.deck_c4(); deck.dt = 0.1
D = 1
lo, hi = (0,), (deck.ncell[0] - 1,)
E, B = decks.analytic_fields(deck, lo, hi, E0=3.0e7, B0=5.0e8)
grid = capi.Grid(D, deck.ncell, deck.xmin, deck.dx, deck.nghost, (1,), volume_scale=deck.volume_scale)
grid.set_fields(E, B)
rng = np.random.default_rng(3)
sdef = deck.species[0]
p = decks.load_species(deck, sdef, lo, hi, rng)
sp = capi.Species(grid, sdef.mass, sdef.charge, sdef.fnorm_const(deck.units), deck.units.cvac_norm, interp_N=deck.interp_N,
                  interp_J=deck.interp_J, interp_E=deck.interp_E, rtol=deck.rtol, iter_max=deck.iter_max)
sp.upload(p["x"], p["v"], p["w"], ids=p["id"])
sp.bin_particles()
def timed(fn, reps=5):
    fn(); capi.check(lib.pgpu_synchronize())
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    capi.check(lib.pgpu_synchronize())
    return (time.perf_counter() - t0) / reps * 1e3
n = sp.n
t_fused = timed(lambda: capi.check(lib.pgpu_advance_particles_iteratively(sp.h, deck.dt, 1, None)))
t_adv = timed(lambda: capi.check(lib.pgpu_advance_particles_iteratively(sp.h, deck.dt, 0, None)))
t_dep = timed(lambda: capi.check(lib.pgpu_set_current_density(sp.h, deck.dt, 0)))
adv, app, unc = capi.picard_totals(reset=True)
print("C4 one species n=%d: fused %.2f ms, advance only %.2f ms, deposit only %.2f ms; passes/particle %.2f" %
      (n, t_fused, t_adv, t_dep, app / max(adv, 1)))
