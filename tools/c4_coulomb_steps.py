"""A few weighted-Coulomb (NANBU, Clog 10) collision steps on a C4-shaped problem (1D, 200 ppc x 2 species) for ncu:
   ncu --set full --clock-control none --import-source on -k regex:'k_coulomb_' -s 3 -c 3 -o gpurun_out/c4c python tools/c4_coulomb_steps.py [ncell]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from picnic_b200 import capi, decks

capi.init(0)
lib = capi.load()
ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 25000
deck = decks.deck_c4(ncell=ncell)
lo, hi = (0,), (deck.ncell[0] - 1,)
grid = capi.Grid(1, deck.ncell, deck.xmin, deck.dx, deck.nghost, (1,), volume_scale=deck.volume_scale)
rng = np.random.default_rng(3)
sps = []
for sdef in deck.species:
    p = decks.load_species(deck, sdef, lo, hi, rng)
    sp = capi.Species(grid, sdef.mass, sdef.charge, sdef.fnorm_const(deck.units), deck.units.cvac_norm,
                      interp_N=deck.interp_N, interp_J=deck.interp_J, interp_E=deck.interp_E)
    sp.upload(p["x"], p["v"], p["w"], ids=p["id"])
    sp.bin_particles(); sp.set_moments()
    sps.append(sp)
grid.debye_length(sps)
dt_sec = 0.1 * deck.units.time
import time
for k in range(3):
    capi.check(lib.pgpu_synchronize()); t0 = time.perf_counter()
    for (a, b) in ((0, 0), (1, 1), (0, 1)):
        capi.check(lib.pgpu_collide_coulomb(sps[a].h, sps[b].h, capi.C.byref(capi.CoulombParams(10.0, 1, 0, 11, 1)),
                                            dt_sec, 1983, k, None))
    capi.check(lib.pgpu_synchronize())
    print("step %d: %.3f ms for %d particles" % (k, (time.perf_counter() - t0) * 1e3, sum(s.n for s in sps)))
