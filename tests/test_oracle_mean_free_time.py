"""Scattering::setMeanFreeTime restatement (oracle/oracle_scatter.cpp) against independent numpy
closed forms of the same reference formulas (TakizukaAbe.cpp:80-238, Coulomb.cpp:108-356,
Elastic.cpp:146-202, MathUtils.cpp:65-95)."""
import math

import numpy as np
from scipy import special

from common import orc

CVAC, ME, QE = 2.99792458e+08, 9.10938370e-31, 1.60217663e-19
EP0 = 1.0 / CVAC / CVAC / (4.0e-7 * math.pi)


def _moments(rng, ncell, mass, T_eV, n0, drift=0.0):
    """Cell moments as set*DensityFromBinFab leaves them: dens, mom = m <n beta>, ene_k = m <n beta_k^2>/2."""
    dens = n0 * rng.uniform(0.5, 1.5, ncell)
    T = T_eV * rng.uniform(0.5, 2.0, ncell)
    vt2 = QE * T / (ME * mass) / CVAC ** 2            # <beta_k^2> thermal
    ub = drift * rng.standard_normal((3, ncell))
    mom = mass * dens * ub
    ene = 0.5 * mass * dens * (vt2[None, :] + ub ** 2)
    return dens, mom, ene


def test_gammainc_series_is_the_lower_incomplete_gamma():
    # the Taylor series of int_0^x t^(1/2) e^(-t) dt; the reference truncates it at 40 terms for x < 10
    for x in (1e-3, 0.1, 0.5, 1.0, 3.0, 6.0):
        want = special.gammainc(1.5, x) * special.gamma(1.5)
        assert abs(orc.gammainc_3half(x) - want) < 1e-9 * max(want, 1e-12), x
    assert orc.gammainc_3half(12.0) == math.gamma(1.5)   # x >= 10: the complete gamma function


def test_ta_intra_is_the_nrl_collision_time():
    rng = np.random.default_rng(1)
    m = _moments(rng, 200, 1.0, 100.0, 1e30)
    m[0][::17] = 0.0                                   # empty cells are skipped
    got = orc.ta_nu_max(m, m, -1.0, -1.0, 1.0, 1.0, 3.0, True)
    dens, _, ene = m
    live = dens > 0
    T = 2.0 / 3.0 * ME * CVAC ** 2 * ene.sum(axis=0)[live] / dens[live] / QE
    tau = 3.44e5 * T ** 1.5 / (dens[live] * 1e-6) / 3.0 * math.sqrt(0.5)
    assert abs(got - np.max(1.0 / tau)) < 1e-12 * got


def test_ta_inter_matches_closed_form_and_is_symmetric():
    rng = np.random.default_rng(2)
    me = _moments(rng, 100, 1.0, 150.0, 1e30)
    mi = _moments(rng, 100, 1836.15, 50.0, 1e30)
    got = orc.ta_nu_max(me, mi, -1.0, 1.0, 1.0, 1836.15, 3.0, False)
    assert got == orc.ta_nu_max(mi, me, 1.0, -1.0, 1836.15, 1.0, 3.0, False)
    e1 = ME * CVAC ** 2 * me[2].sum(axis=0) / me[0]
    e2 = ME * CVAC ** 2 * mi[2].sum(axis=0) / mi[0]
    T1, T2 = 2.0 / 3.0 * e1 / QE, 2.0 / 3.0 * e2 / QE
    VT1, VT2 = np.sqrt(QE * T1 / ME), np.sqrt(QE * T2 / (ME * 1836.15))
    x12 = T1 / (T2 / 1836.15)
    G = lambda x: np.where(x < 10, special.gammainc(1.5, np.minimum(x, 10)) * special.gamma(1.5), special.gamma(1.5))
    factor = (QE * QE / EP0) ** 2 / (4 * math.pi)
    nu12 = (1 + 1 / 1836.15) * 2 / math.sqrt(math.pi) * G(x12) * factor * 3.0 * mi[0] / e1 ** 2 * VT1
    nu21 = (1 + 1836.15) * 2 / math.sqrt(math.pi) * G(1 / x12) * factor * 3.0 * me[0] / e2 ** 2 * VT2
    want = max(nu12.max(), nu21.max())
    assert abs(got - want) < 1e-8 * want


def test_coulomb_fixed_clog_and_sigma_cap():
    rng = np.random.default_rng(3)
    m = _moments(rng, 64, 1.0, 100.0, 1e30, drift=1e-3)
    LDe = np.full(64, 1e-9)
    got = orc.coulomb_nu_max(LDe, m, m, -1.0, -1.0, 1.0, 1.0, 10.0, True)
    dens, mom, ene = m
    meanE = (mom ** 2).sum(axis=0) / dens / 2.0
    T = np.maximum(2.0 / 3.0 * (ene.sum(axis=0) - meanE) / dens * ME * CVAC ** 2 / QE, 0.01)
    g2 = 6.0 * QE / ME * T
    hbar = 6.62607015e-34 / (2 * math.pi)
    mu = 0.5
    EF = hbar ** 2 / (2 * ME * mu) * (3 * math.pi ** 2) ** (2.0 / 3.0) / (ME * CVAC ** 2) * dens ** (2.0 / 3.0)
    b90 = (QE * QE / CVAC ** 2 / (2 * math.pi * EP0 * ME)) / (mu * g2 / CVAC ** 2 + 2 * EF)
    smax = 1.0 / (dens / np.cbrt(4.0 / 3.0 * math.pi * dens))
    nu = np.sqrt(g2) * dens * np.minimum(8 / math.pi * b90 ** 2 * 10.0, smax)
    assert abs(got - nu.max()) < 1e-11 * got
    # computed Clog (Lee-More) is floored at 2: a tiny Debye length gives exactly the Clog = 2 answer
    assert orc.coulomb_nu_max(np.full(64, 1e-14), m, m, -1.0, -1.0, 1.0, 1.0, 0.0, True) == \
        orc.coulomb_nu_max(LDe, m, m, -1.0, -1.0, 1.0, 1.0, 2.0, True)


def test_elastic_constant_sigma():
    rng = np.random.default_rng(4)
    me = _moments(rng, 50, 1.0, 5.0, 1e22)
    mn = _moments(rng, 50, 7294.3, 0.03, 3e22)
    got = orc.elastic_nu_max(me, mn, 1.0, 7294.3, const_sigma=6e-20)
    g12 = np.sqrt(2 * me[2].sum(axis=0) / me[0] + 2 * mn[2].sum(axis=0) / (mn[0] * 7294.3))
    assert abs(got - np.max(mn[0] * 6e-20 * g12 * CVAC)) < 1e-13 * got
