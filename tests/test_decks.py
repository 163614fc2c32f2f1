"""CPU tests of the synthetic deck builder (host logic only)."""
import numpy as np

from picnic_b200 import decks
from common import orc


def test_loader_lattice_and_weights():
    deck = decks.deck_c2(ncell=4, ppc=3)
    rng = np.random.default_rng(0)
    p = decks.load_species(deck, deck.species[0], (0, 0), (3, 3), rng)
    n = 4 * 4 * 9
    assert p["x"].shape == (2, n) and p["v"].shape == (3, n)
    # sub-cell lattice: x = x_cell_lo + (i+1/2)*dx/ppc  (PicChargedSpecies.cpp:2346-2360)
    frac = (p["x"][0] / 0.25) % 1.0
    assert np.allclose(np.unique(np.round(frac, 12)), [1 / 6, 0.5, 5 / 6])
    # cell ordered: bins are non-decreasing
    g = orc.make_geom(2, deck.xmin, deck.xmax, deck.dx, 2)
    c = orc.bin_cells(g, p["x"])
    lin = c[0] + 4 * c[1]
    assert np.all(np.diff(lin) >= 0) and np.all(np.bincount(lin) == 9)
    # w = n*dV/ppc with dV in m^3
    assert np.allclose(p["w"], 1e30 * (0.25 * 5.314e-9) ** 2 / 9)


def test_maxwellian_temperature():
    deck = decks.deck_c2(ncell=16, ppc=8)
    rng = np.random.default_rng(1)
    e = decks.load_species(deck, deck.species[0], (0, 0), (15, 15), rng)
    # T_eV = m c^2 <beta^2> / qe
    T = decks.ME * decks.CVAC ** 2 * (e["v"] ** 2).mean(axis=1) / decks.QE
    assert np.allclose(T, 150.0, rtol=0.03)


def test_fields_are_periodic_and_staggered():
    deck = decks.deck_c3(ncell=8, ppc=2)
    E, B = decks.analytic_fields(deck, (0, 0), (7, 7))
    lo, hi, ex = E[0]
    assert lo == (-3, -3) and hi == (10, 11)        # Ex: cell-centred in x, nodal in y
    g = deck.nghost
    assert np.allclose(ex[g:g + 8, g], ex[g:g + 8, g + 8], atol=1e-6 * np.abs(ex).max())
    lo, hi, bz = B[2]
    assert hi == (10, 10)
    assert np.allclose(bz[0, :], bz[8, :], atol=1e-6 * np.abs(bz).max())
