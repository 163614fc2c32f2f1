"""A few C2 collision steps (prepForScatter + e-e, i-i, e-i Takizuka-Abe) for profiling under ncu:
   ncu --set full --clock-control none --import-source on -k regex:'k_ta_|k_bin_|k_cell' -c 24 -o gpurun_out/c2 python tools/c2_steps.py"""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
from picnic_b200 import capi as pgpu, decks

pgpu.load(); pgpu.init(0)
d = decks.deck_c2()
lo, hi = (0, 0), (d.ncell[0] - 1, d.ncell[1] - 1)
grid = pgpu.Grid(2, d.ncell, d.xmin, d.dx, d.nghost, (1, 1), volume_scale=d.volume_scale)
rng = np.random.default_rng(d.seed)
sps = []
for sdef in d.species:
    p = decks.load_species(d, sdef, lo, hi, rng)
    sp = pgpu.Species(grid, sdef.mass, sdef.charge, sdef.fnorm_const(d.units), d.units.cvac_norm,
                      interp_N=d.interp_N, interp_J=d.interp_J, interp_E=d.interp_E)
    sp.upload(p["x"], p["v"], p["w"], ids=p["id"])
    sps.append(sp)
dt_sec = d.dt * d.units.time
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for step in range(nsteps):
    for sp in sps:
        sp.update_old_positions(); sp.update_old_velocities()
        sp.bin_particles(); sp.set_moments()
    for (a, b) in ((0, 0), (1, 1), (0, 1)):
        pgpu.check(pgpu.load().pgpu_collide_ta(sps[a].h, sps[b].h, 3.0, dt_sec, 1983, step, None))
pgpu.check(pgpu.load().pgpu_synchronize())
print("done")
