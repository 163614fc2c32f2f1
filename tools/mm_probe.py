#!/usr/bin/env python
"""Device timings of the mass-matrix path on the C3 box (2D 512^2, 2 species x ppc^2 particles per cell, CC1):
setMassMatrices (zero + accumulate both species + save E0) and computeJfromMassMatrices, run kernel vs the
generic one-thread-per-particle kernel.  Not a bench line; DESIGN.md quotes the numbers."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from picnic_b200 import capi, decks  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ncell", type=int, default=512)
ap.add_argument("--ppc", type=int, default=10)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--sort", default="dual", choices=["dual", "cell"])
args = ap.parse_args()

capi.init(0)
lib = capi.load()
deck = decks.deck_c3(ncell=args.ncell, ppc=args.ppc, dt=0.1)
lo, hi = (0, 0), tuple(n - 1 for n in deck.ncell)
E, B = decks.analytic_fields(deck, lo, hi, E0=3.0e7, B0=5.0e8)
grid = capi.Grid(2, deck.ncell, deck.xmin, deck.dx, deck.nghost, (1, 1), volume_scale=deck.volume_scale)
grid.set_fields(E, B)
rng = np.random.default_rng(3)
sps = []
for sdef in deck.species:
    p = decks.load_species(deck, sdef, lo, hi, rng)
    sp = capi.Species(grid, sdef.mass, sdef.charge, sdef.fnorm_const(deck.units), deck.units.cvac_norm,
                      interp_N=deck.interp_N, interp_J=deck.interp_J, interp_E=deck.interp_E, rtol=deck.rtol,
                      iter_max=deck.iter_max)
    sp.upload(p["x"], p["v"], p["w"], ids=p["id"])
    sp.sort_for_locality() if args.sort == "dual" else sp.bin_particles()
    sps.append(sp)
n = sum(sp.n for sp in sps)
# one implicit advance so that (xbar, ubar) are a converged orbit
for sp in sps:
    capi.check(lib.pgpu_advance_particles_iteratively(sp.h, deck.dt, 1, None))
nc = grid.mass_matrices_init(capi.CC1 if hasattr(capi, "CC1") else 3)
ncomp = int(sum(int(a) * int(b) for a, b in nc))
print("particles %d, sigma components %d (%.2f GB)" % (n, ncomp, ncomp * (args.ncell + 7) ** 2 * 8 / 1e9))


def timed(fn, reps):
    fn(); capi.check(lib.pgpu_synchronize())
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    capi.check(lib.pgpu_synchronize())
    return (time.perf_counter() - t0) / reps


def set_mm():
    grid.mass_matrices_zero()
    for sp in sps:
        sp.accumulate_mass_matrices(deck.dt)
    grid.mass_matrices_save_E0()


for mode, name in ((1, "run kernel"), (0, "generic kernel")):
    lib.pgpu_set_deposit_mode(mode)
    set_mm(); capi.check(lib.pgpu_synchronize())      # first launches stay out of the kernel timers
    lib.pgpu_profile_enable(1)
    lib.pgpu_profile_reset()
    t = timed(set_mm, args.reps if mode else 1)
    ms, k = capi.C.c_double(), capi.C.c_long()
    parts = {}
    for pre in ("mass_matrix_run", "mass_matrix_deferred", "mass_matrix_generic"):
        lib.pgpu_profile_query(pre.encode(), capi.C.byref(ms), capi.C.byref(k))
        if k.value:
            parts[pre] = ms.value / k.value
    lib.pgpu_profile_enable(0)
    print("setMassMatrices (%s): %.2f ms = %.3e particles/s; per launch [ms]: %s" % (
        name, t * 1e3, n / t, {a: round(b, 3) for a, b in parts.items()}))
lib.pgpu_set_deposit_mode(1)
capi.check(lib.pgpu_picard_totals(None, None, None, 1))
t = timed(grid.compute_J_from_mass_matrices, args.reps)
print("computeJfromMassMatrices: %.3f ms (%.0f GB/s of sigma)" % (t * 1e3, ncomp * (args.ncell + 7) ** 2 * 8 / t / 1e9))
# sanity: J from the matrices at E == E0 is J0 == the deposited current of the same orbits
grid.current_zero()
for sp in sps:
    grid.current_add(sp)
Jdep = [grid.current_get(c) for c in range(3)]
grid.compute_J_from_mass_matrices()
for c in range(3):
    J = grid.current_get(c)
    print("J0 vs fused deposit, comp %d: %.2e" % (c, np.max(np.abs(J - Jdep[c])) / np.max(np.abs(Jdep[c]))))
capi.finalize()
