import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from picnic_b200 import capi as pgpu, decks
from oracle import oracle as orc
pgpu.load(); pgpu.init(0)
deck = decks.deck_c2()
Clog = 3.0
def run_gpu(ncell, dtf, nsteps, seed):
    d = decks.deck_c2(ncell=ncell)
    lo, hi = (0, 0), (ncell - 1, ncell - 1)
    grid = pgpu.Grid(2, d.ncell, d.xmin, d.dx, d.nghost, (1, 1), volume_scale=d.volume_scale)
    rng = np.random.default_rng(12)
    sps = []
    for sdef in d.species:
        p = decks.load_species(d, sdef, lo, hi, rng)
        sp = pgpu.Species(grid, sdef.mass, sdef.charge, sdef.fnorm_const(d.units), d.units.cvac_norm)
        sp.upload(p["x"], p["v"], p["w"], ids=p["id"])
        sp.bin_particles(); sp.set_moments()
        sps.append(sp)
    dt_sec = dtf * d.dt * d.units.time
    mass = [s.mass for s in d.species]
    def temps():
        return [mk * sp.global_moments()[4:7].sum() / sp.global_moments()[0] for mk, sp in zip(mass, sps)]
    T0 = temps()
    hist = []
    for step in range(nsteps):
        for sp in sps:
            sp.bin_particles(); sp.set_moments()
        for (a, b) in ((0, 0), (1, 1), (0, 1)):
            pgpu.collide_ta(sps[a], sps[b], Clog, dt_sec, seed, step, count=False)
        hist.append(temps())
    for sp in sps: sp.destroy()
    grid.destroy()
    return T0, hist, dt_sec
def run_cpu(ncell, dtf, nsteps, seed):
    d = decks.deck_c2(ncell=ncell)
    lo, hi = (0, 0), (ncell - 1, ncell - 1)
    rng = np.random.default_rng(12)
    ps = [decks.load_species(d, sdef, lo, hi, rng) for sdef in d.species]
    nc = ncell * ncell
    cs = np.arange(nc + 1, dtype=np.int64) * 64
    cellV = 0.25 * 0.25 * d.volume_scale
    dens = [np.full(nc, p["w"][:64].sum() / cellV) for p in ps]
    dt_sec = dtf * d.dt * d.units.time
    mass = [s.mass for s in d.species]; q = [s.charge for s in d.species]
    v = [p["v"].copy() for p in ps]
    def temps(): return [m * (vv ** 2).sum() / vv.shape[1] for m, vv in zip(mass, v)]
    T0 = temps(); hist = []
    orc.lib().orc_rng_seed(seed)
    for step in range(nsteps):
        orc.ta_self(cs, v[0], dens[0], mass[0], q[0], Clog, dt_sec)
        orc.ta_self(cs, v[1], dens[1], mass[1], q[1], Clog, dt_sec)
        orc.ta_inter(cs, v[0], dens[0], mass[0], q[0], cs, v[1], dens[1], mass[1], q[1], Clog, dt_sec)
        hist.append(temps())
    return T0, hist, dt_sec
for dtf, nsteps in ((10, 40), (40, 40)):
    g, c = [], []
    for seed in (1983, 7, 21, 99):
        T0, h, dt_sec = run_gpu(64, dtf, nsteps, seed)
        dT0 = T0[0] - T0[1]; g.append((dT0 - (h[-1][0] - h[-1][1])) / dT0)
        T0, h, dt_sec = run_cpu(64, dtf, nsteps, seed)
        dT0 = T0[0] - T0[1]; c.append((dT0 - (h[-1][0] - h[-1][1])) / dT0)
    print(dtf, "GPU", g, np.mean(g), "CPU", c, np.mean(c), "ratio", np.mean(g) / np.mean(c), flush=True)
