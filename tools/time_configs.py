#!/usr/bin/env python
"""Device timings of the secondary BASELINE configurations (not bench lines):
C4  1D 250 000 cells x 200 ppc x 2 species: fused implicit advance+deposit, weighted Coulomb (NANBU), cell sort
C1  1D 40 cells x 100 ppc x 2 species (launch-latency bound)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from picnic_b200 import capi, decks

capi.init(0)
lib = capi.load()


def timed(fn, reps=5):
    fn(); capi.check(lib.pgpu_synchronize())
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    capi.check(lib.pgpu_synchronize())
    return (time.perf_counter() - t0) / reps


def run(deck, name, ncell_scale=1):
    D = deck.D
    lo, hi = tuple([0] * D), tuple(n - 1 for n in deck.ncell)
    E, B = decks.analytic_fields(deck, lo, hi, E0=3.0e7, B0=5.0e8)
    grid = capi.Grid(D, deck.ncell, deck.xmin, deck.dx, deck.nghost, tuple([1] * D), volume_scale=deck.volume_scale)
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
    n = sum(sp.n for sp in sps)
    def adv():
        for sp in sps:
            capi.check(lib.pgpu_advance_particles_iteratively(sp.h, deck.dt, 1, None))
    t_adv = timed(adv)
    def srt():
        for sp in sps:
            sp.bin_particles(); sp.set_moments()
    t_sort = timed(srt)
    capi.profile_reset(); capi.profile_enable(True)
    srt(); capi.check(lib.pgpu_synchronize())
    capi.profile_enable(False)
    print("  sort breakdown (ms): " + ", ".join("%s %.3f" % (k, capi.profile_query(k)[0])
                                                for k in ("bin_key", "bin_sort", "bin_permute", "bin_starts", "cell_moments")))
    grid.debye_length(sps)
    dt_sec = deck.dt * deck.units.time
    state = {"k": 0}
    def col():
        for (a, b) in ((0, 0), (1, 1), (0, 1)):
            capi.check(lib.pgpu_collide_coulomb(sps[a].h, sps[b].h, capi.C.byref(capi.CoulombParams(10.0, 1, 0, 11, 1)),
                                                dt_sec, 1983, state["k"], None))
        state["k"] += 1
    t_col = timed(col)
    pairs = sum(capi.collide_coulomb(sps[a], sps[b], 10.0, dt_sec, 1983, 99, angular=1) for (a, b) in ((0, 0), (1, 1), (0, 1)))
    print("%s: %d particles | advance+deposit %.3f ms = %.3e advances/s (%.1f%% of HBM roofline at %d B) | sort+moments %.3f ms | "
          "Coulomb NANBU %.3f ms = %.3e pairs/s" % (name, n, t_adv * 1e3, n / t_adv,
          100 * n * (2 * D + 7) * 8 / t_adv / 6.5443e12, (2 * D + 7) * 8, t_sort * 1e3, t_col * 1e3, pairs / t_col))
    for sp in sps:
        sp.destroy()
    grid.destroy()


d4 = decks.deck_c4(); d4.dt = 0.1
run(d4, "C4")
d1 = decks.deck_c1(); d1.iter_max = 21
run(d1, "C1")
capi.finalize()
