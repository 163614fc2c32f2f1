"""Synthetic input decks for the per-particle hot path.

Uniform two-species Maxwellian plasma, cartesian, periodic, in the units of the
reference deck regression_tests/2d/numericalEnergy_theta_implicit/thermalization.in
(units.length = 5.314e-9 m, units.time = 1.77e-17 s, n = 1e30 m^-3).  Particles are
laid out the way PICNIC's loader does it (PicChargedSpecies.cpp:2328-2376):
a uniform sub-cell lattice, Gaussian velocities beta = sqrt(qe/me*T/m)*N(0,1)/c,
equal weights w = n*dV/ppc.  Grid fields are smooth analytic periodic modes
sampled at each component's Yee location, ghosts included.

Everything here is host-side numpy; it produces inputs, it is not on the hot path.
"""
import math
from dataclasses import dataclass, field

import numpy as np

# src/core/PicnicConstants.H:13-51
PI = math.pi
CVAC = 2.99792458e+08
MU0 = 4.0 * PI * 1.0e-7
EP0 = 1.0 / CVAC / CVAC / MU0
ME = 9.10938370e-31
QE = 1.60217663e-19

CIC, TSC, CC0, CC1 = 0, 1, 2, 3
INTERP_NAMES = {"CIC": CIC, "TSC": TSC, "CC0": CC0, "CC1": CC1}

E_STAG = {1: [(0,), (1,), (1,)], 2: [(0, 1), (1, 0), (1, 1)]}
B_STAG = {1: [(1,), (0,), (0,)], 2: [(1, 0), (0, 1), (0, 0)]}


@dataclass
class Units:
    """CodeUnits (src/core/CodeUnits.cpp:18-40)."""
    length: float = 5.314e-9
    time: float = 1.77e-17
    number_density: float = 1.0

    @property
    def cvac_norm(self):
        return CVAC * self.time / self.length

    @property
    def escale(self):
        return 1.0e5


@dataclass
class SpeciesDef:
    name: str
    mass: float
    charge: float
    temperature_eV: tuple = (100.0, 100.0, 100.0)
    density: float = 1.0e30
    ppc: tuple = (10, 10)   # particles per cell per direction

    def fnorm_const(self, units):
        """m_fnorm_const (PicChargedSpecies.cpp:1953-1958)."""
        qom = self.charge / self.mass * QE / ME
        cvacSq = CVAC * CVAC
        return qom / cvacSq * units.escale * units.length


@dataclass
class Deck:
    D: int
    ncell: tuple
    dx: tuple
    xmin: tuple
    nghost: int
    units: Units = field(default_factory=Units)
    dt: float = 0.1
    species: list = field(default_factory=list)
    interp_E: int = CC1
    interp_J: int = CC1
    interp_N: int = TSC
    rtol: float = 1.0e-12
    iter_max: int = 21
    seed: int = 1983

    @property
    def xmax(self):
        return tuple(x0 + n * h for x0, n, h in zip(self.xmin, self.ncell, self.dx))

    @property
    def cnorm_dt(self):
        return self.dt * self.units.cvac_norm

    @property
    def volume_scale(self):
        """DomainGrid m_volume_scale for cartesian geometry (DomainGrid.cpp:137-138)."""
        return self.units.length ** self.D


def electron_proton(ppc, Te=100.0, Ti=100.0, density=1.0e30):
    return [SpeciesDef("electron", 1.0, -1.0, (Te,) * 3, density, ppc),
            SpeciesDef("proton", 1836.15, 1.0, (Ti,) * 3, density, ppc)]


def fab_shape(box_lo, box_hi, nghost, stag):
    lo = tuple(l - nghost for l in box_lo)
    hi = tuple(h + nghost + s for h, s in zip(box_hi, stag))
    return lo, hi


def load_species(deck, sp, box_lo, box_hi, rng, dtype=np.float64):
    """Lattice positions + Maxwellian velocities for the cells box_lo..box_hi.

    Returns dict with x[D,n], v[3,n], w[n], id[n] (cell-ordered, dir 0 fastest,
    so the arrays are already cell sorted)."""
    D = deck.D
    ncell_box = [h - l + 1 for l, h in zip(box_lo, box_hi)]
    ppc = sp.ppc[:D]
    ppc_tot = int(np.prod(ppc))
    ncells = int(np.prod(ncell_box))
    n = ncells * ppc_tot
    x = np.empty((D, n), dtype=dtype)
    if D == 1:
        ic = np.arange(box_lo[0], box_hi[0] + 1, dtype=np.float64)
        sub = (np.arange(ppc[0]) + 0.5) * (deck.dx[0] / ppc[0])
        # Xpart = Xcc - 0.5*dX + (ipg+0.5)*dXpart ; Xcc = xmin + (i+0.5)*dx
        xcc = deck.xmin[0] + (ic + 0.5) * deck.dx[0]
        x[0] = ((xcc - 0.5 * deck.dx[0])[:, None] + sub[None, :]).reshape(-1)
    else:
        i = np.arange(box_lo[0], box_hi[0] + 1, dtype=np.float64)
        j = np.arange(box_lo[1], box_hi[1] + 1, dtype=np.float64)
        s0 = (np.arange(ppc[0]) + 0.5) * (deck.dx[0] / ppc[0])
        s1 = (np.arange(ppc[1]) + 0.5) * (deck.dx[1] / ppc[1])
        x0c = (deck.xmin[0] + (i + 0.5) * deck.dx[0]) - 0.5 * deck.dx[0]
        x1c = (deck.xmin[1] + (j + 0.5) * deck.dx[1]) - 0.5 * deck.dx[1]
        # order: cell j (slow), cell i, sub j, sub i (fast)
        X0 = x0c[None, :, None, None] + s0[None, None, None, :]
        X1 = x1c[:, None, None, None] + s1[None, None, :, None]
        shape = (ncell_box[1], ncell_box[0], ppc[1], ppc[0])
        x[0] = np.broadcast_to(X0, shape).reshape(-1)
        x[1] = np.broadcast_to(X1, shape).reshape(-1)
    V0 = math.sqrt(QE / ME)
    v = np.empty((3, n), dtype=dtype)
    for c in range(3):
        vt = V0 * math.sqrt(sp.temperature_eV[c] / sp.mass)
        v[c] = vt * rng.standard_normal(n) / CVAC
    cell_volume = deck.volume_scale * float(np.prod(deck.dx[:D]))
    pweight = deck.units.number_density * sp.density * cell_volume / float(ppc_tot)
    w = np.full(n, pweight, dtype=dtype)
    ids = np.arange(n, dtype=np.uint64)
    return {"x": x, "v": v, "w": w, "id": ids}


def analytic_fields(deck, box_lo, box_hi, E0=2.0e7, B0=4.0e8, modes=3, kmul=1):
    """Six field components on the ghosted arrays of the box, Yee-staggered.

    Smooth periodic sums of `modes` low-k modes (wave numbers kmul * 2 pi (m+1) / L: kmul > 1 shortens the
    wavelengths, the Picard particle loop then needs more passes).  Returns (E, B, meta) with
    E[c], B[c] = (lo, hi, array(F-order))."""
    D = deck.D
    L = [n * h for n, h in zip(deck.ncell, deck.dx)]
    rng = np.random.default_rng(deck.seed + 77)
    ph = rng.uniform(0, 2 * PI, size=(6, modes, 2))
    amp = rng.uniform(0.5, 1.0, size=(6, modes))

    def coords(stag, d, lo, hi):
        idx = np.arange(lo, hi + 1, dtype=np.float64)
        return deck.xmin[d] + (idx + (0.0 if stag else 0.5)) * deck.dx[d]

    def comp(ci, stag, scale):
        lo, hi = fab_shape(box_lo, box_hi, deck.nghost, stag)
        X = coords(stag[0], 0, lo[0], hi[0])
        if D == 1:
            a = np.zeros(X.shape)
            for m in range(modes):
                k = 2 * PI * kmul * (m + 1) / L[0]
                a += amp[ci, m] * np.sin(k * (X - deck.xmin[0]) + ph[ci, m, 0])
        else:
            Y = coords(stag[1], 1, lo[1], hi[1])
            a = np.zeros((X.size, Y.size))
            for m in range(modes):
                kx = 2 * PI * kmul * (m + 1) / L[0]
                ky = 2 * PI * kmul * (modes - m) / L[1]
                a += amp[ci, m] * np.outer(np.sin(kx * (X - deck.xmin[0]) + ph[ci, m, 0]),
                                           np.cos(ky * (Y - deck.xmin[1]) + ph[ci, m, 1]))
        return lo, hi, np.asfortranarray(scale * a / modes)

    E = [comp(c, E_STAG[D][c], E0) for c in range(3)]
    B = [comp(3 + c, B_STAG[D][c], B0) for c in range(3)]
    return E, B


# ---- named configurations (BASELINE.json configs / BASELINE.md table) ----------

def deck_c1():
    """C1: 1D 40 cells x 100 ppc x 2 species, CC1/CC1/TSC, iter_max_particles=0."""
    d = Deck(D=1, ncell=(40,), dx=(0.25,), xmin=(0.0,), nghost=2, dt=0.1, iter_max=0)
    d.species = electron_proton((100,))
    return d


def deck_c2(ncell=256, ppc=8):
    """C2: 2D 256x256, 64 ppc/species, TA collisions (Te=150eV, Ti=50eV variant)."""
    d = Deck(D=2, ncell=(ncell, ncell), dx=(0.25, 0.25), xmin=(0.0, 0.0), nghost=2, dt=0.1)
    d.species = electron_proton((ppc, ppc), Te=150.0, Ti=50.0)
    return d


def deck_c3(ncell=512, ppc=10, dt=3.0, iter_max=21):
    """C3: 2D 512x512, 2 species x 100 ppc, CC1/CC1/TSC, rtol 1e-12."""
    d = Deck(D=2, ncell=(ncell, ncell), dx=(0.25, 0.25), xmin=(0.0, 0.0), nghost=3, dt=dt,
             iter_max=iter_max)
    d.species = electron_proton((ppc, ppc))
    return d


def deck_c4(ncell=250000, ppc=200):
    """C4: 1D shock stand-in, 2 species x ncell x 200 ppc (periodic for timing)."""
    d = Deck(D=1, ncell=(ncell,), dx=(0.25,), xmin=(0.0,), nghost=3, dt=3.0)
    # positions reach 6e4 code units, where one ulp is 3e-11 dx: the step-norm test |dxp0 - dxp|/dx
    # cannot resolve the 1e-12 of the small decks (the reference's long 1D decks use 1e-8 .. 1e-10)
    d.rtol = 1.0e-8
    d.species = electron_proton((ppc,))
    return d
