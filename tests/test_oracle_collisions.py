"""CPU checks of the Coulomb / Elastic restatement in the oracle (reference-derived properties):
Nanbu's A(s) fit reproduces <cos theta> = exp(-s) (Nanbu 1997; the relation the reference quotes in
Coulomb.H:343-344), equal-weight Coulomb pairs conserve momentum and energy, the pair counts are the
reference's (NxN below NxN_Nthresh, odd-N triple), the weighted rejection conserves momentum and
energy on average, and the elastic cross-section lookup interpolates as the reference writes it."""
import numpy as np
import pytest

from common import orc

DT_SEC = 0.1 * 1.77e-17


def test_nanbu_mean_cosine():
    U = (np.arange(20000) + 0.5) / 20000
    for s in (0.01, 0.1, 0.3, 1.0, 2.5, 4.0, 8.0):
        c = np.array([orc.nanbu_costh_sinth(s, u)[0] for u in U])
        assert np.all(np.abs(c) <= 1.0 + 1e-12)
        assert abs(c.mean() - np.exp(-s)) < (0.012 if s < 6 else 1e-3), s      # s >= 6: isotropic


def _cells(rng, counts):
    cs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    n = int(cs[-1])
    v = rng.standard_normal((3, n)) * 0.02
    return cs, v


@pytest.mark.parametrize("variant", [3, 4])
def test_nanbu_full_angle_scattering_restatement(variant):
    """NANBU_FAS / NANBU_FAS_v2 (Coulomb.H:365-718): the cumulative branch above the switch-over is Nanbu's model with the
    same draw; the mean of 1 - cos(theta) follows what the model is built for -- Nanbu's 1 - exp(-s12) for FAS; for v2
    the same minus the part of the Rutherford tail its sampling formula leaves out,
    s12/(2 Clog) mu_max/(mu_max + mu_tr) with mu_tr = s12/Clog; draws are consumed in order and only when needed."""
    rng = np.random.default_rng(17 + variant)
    Clog, b0 = 5.0, 1.0e-9
    bmin = 0.15 * b0
    sigma_eff = np.pi * b0 * b0 * Clog
    above = 0.5 if variant == 3 else 0.6000001
    for s12 in (above, 1.0, 4.0, 7.0):
        for u in rng.random(20):
            c, sn = orc.nanbu_fas_costh_sinth(variant, s12, Clog, b0, bmin, sigma_eff, u, 0.123, 0.987)
            assert (c, sn) == orc.nanbu_costh_sinth(s12, u)
    bperp_sq, bmin_sq = b0 * b0 / 4.0, bmin * bmin
    mu_max = bperp_sq / (bmin_sq + bperp_sq)
    for s12 in (0.05, 0.2, 0.45):
        U = rng.random((40000, 3))
        c = np.array([orc.nanbu_fas_costh_sinth(variant, s12, Clog, b0, bmin, sigma_eff, *u)[0] for u in U])
        assert np.all(np.abs(c) <= 1.0)
        want = 1.0 - np.exp(-s12)
        if variant == 4:
            mu_tr = s12 / Clog
            want -= s12 / (2.0 * Clog) * mu_max / (mu_max + mu_tr)
        assert abs(np.mean(1.0 - c) - want) / want < 0.03, (s12, np.mean(1.0 - c), want)
    # far below the switch-over: at most one Rutherford event, probability ~ N12, else no deflection
    s12 = 1.0e-6
    U = rng.random((20000, 3))
    c = np.array([orc.nanbu_fas_costh_sinth(variant, s12, Clog, b0, bmin, sigma_eff, *u)[0] for u in U])
    assert 0.0 < np.mean(c != 1.0) < 0.05
    # the second draw is only looked at when the first one asks for it
    c1 = orc.nanbu_fas_costh_sinth(variant, s12, Clog, b0, bmin, sigma_eff, 0.999999, 0.1, 0.2)
    c2 = orc.nanbu_fas_costh_sinth(variant, s12, Clog, b0, bmin, sigma_eff, 0.999999, 0.9, 0.8)
    if variant == 3:
        assert c1 == c2 == (1.0, 0.0)          # rand() < PL fails: no event, RL never drawn


@pytest.mark.parametrize("angular", [0, 1, 2, 3, 4, 5])
def test_coulomb_intra_equal_weights_conserve_and_count(angular):
    rng = np.random.default_rng(3)
    counts = np.array([0, 1, 2, 3, 5, 10, 11, 12, 13, 40, 41])
    cs, v = _cells(rng, counts)
    n = v.shape[1]
    w = np.full(n, 2.0e27)
    cellV = 1.0e-3
    dens = counts * 2.0e27 / cellV
    LDe = np.full(counts.size, 1.0e-9)
    v0 = v.copy()
    orc.lib().orc_rng_seed(5)
    npairs = orc.coulomb_intra(cs, v, w, dens, LDe, cellV, 1.0, -1.0, 0.0, angular, False, 11, DT_SEC)

    def expect(c):
        if c < 2:
            return 0
        if c < 11:
            return c * (c - 1) // 2
        return c // 2 if c % 2 == 0 else (c - 3) // 2 + 3
    assert npairs == sum(expect(c) for c in counts)
    for k, c in enumerate(counts):
        a, b = cs[k], cs[k + 1]
        if c < 2:
            assert np.array_equal(v[:, a:b], v0[:, a:b])
            continue
        assert np.max(np.abs(v[:, a:b].sum(axis=1) - v0[:, a:b].sum(axis=1))) < 1e-15 * c
        assert abs((v[:, a:b] ** 2).sum() - (v0[:, a:b] ** 2).sum()) / (v0[:, a:b] ** 2).sum() < 1e-13


def test_coulomb_large_angle_scattering_keeps_the_variance_and_conserves():
    """include_large_angle_scattering (Coulomb.cpp:1801-1863): single Rutherford events below a cutoff impact parameter, the
    small-angle s12 reduced "in order to keep total variance equal to s12".  (a) every pair update is a rotation of the
    relative velocity; (b) the mean of 1 - cos(theta) over the event draw and the polar draw stays within a few per cent of
    the small-angle model's; (c) the cell driver draws the extra uniform and still conserves momentum and energy."""
    rng = np.random.default_rng(11)
    v1, v2 = np.array([0.03, -0.01, 0.02]), np.array([-0.02, 0.015, 0.0])
    u = np.linalg.norm(v1 - v2)
    args = dict(EF=1.0e-7, Clog=3.0, den=1.0e30, bmax=1.0e-9, smax=1.0e-15)   # N12 ~ 10: one pair in ten makes an event

    def mean_one_minus_cos(on, ndraw=20000):
        acc, events = 0.0, 0
        for _ in range(ndraw):
            orc.coulomb_set_large_angle(on, rng.random())
            dU, s12 = orc.coulomb_delta_u(v1, v2, -1.0, 1.0, 1.0, 1.0, args["EF"], args["Clog"], 1, args["den"], args["bmax"],
                                          args["smax"], 4.0 * DT_SEC, 0.0, rng.random(), rng.random())
            un = v1 - v2 + dU
            assert abs(np.linalg.norm(un) - u) / u < 1e-12
            acc += 1.0 - np.dot(un, v1 - v2) / u ** 2
            events += s12 < 0
        return acc / ndraw, events
    try:
        m_off, e_off = mean_one_minus_cos(False)
        m_on, e_on = mean_one_minus_cos(True)
        assert e_off == 0 and e_on > 100
        assert abs(m_on - m_off) / m_off < 0.08
        counts = np.array([0, 1, 2, 5, 11, 12, 13, 40])
        cs, v = _cells(rng, counts)
        w = np.full(v.shape[1], 2.0e27)
        cellV = 1.0e-3
        v0 = v.copy()
        orc.coulomb_set_large_angle(True)
        orc.lib().orc_rng_seed(5)
        orc.coulomb_intra(cs, v, w, counts * 2.0e27 / cellV, np.full(counts.size, 1.0e-9), cellV, 1.0, -1.0, 0.0, 1, False, 11,
                          DT_SEC)
        assert np.abs(v - v0).max() > 0
        for k, c in enumerate(counts):
            a, b = cs[k], cs[k + 1]
            if c >= 2:
                assert np.max(np.abs(v[:, a:b].sum(axis=1) - v0[:, a:b].sum(axis=1))) < 1e-15 * c
                assert abs((v[:, a:b] ** 2).sum() - (v0[:, a:b] ** 2).sum()) / (v0[:, a:b] ** 2).sum() < 1e-13
    finally:
        orc.coulomb_set_large_angle(False)


def test_coulomb_inter_weighted_conserves_on_average():
    """Unequal weights: the lighter-weight particle always scatters, the heavier with probability
    wmin/wmax, so momentum and energy are conserved in expectation (Coulomb.cpp:1149-1172)."""
    rng = np.random.default_rng(8)
    ncell, n1c, n2c = 400, 24, 16
    cs1 = np.arange(ncell + 1, dtype=np.int64) * n1c
    cs2 = np.arange(ncell + 1, dtype=np.int64) * n2c
    v1 = rng.standard_normal((3, ncell * n1c)) * 0.02
    v2 = rng.standard_normal((3, ncell * n2c)) * 0.0008
    v1[0] += 0.01
    w1 = np.where(rng.random(ncell * n1c) < 0.5, 1.0e27, 3.0e27)
    w2 = np.where(rng.random(ncell * n2c) < 0.5, 2.0e27, 0.5e27)
    cellV = 1.0e-3
    dens1 = np.add.reduceat(w1, cs1[:-1]) / cellV
    dens2 = np.add.reduceat(w2, cs2[:-1]) / cellV
    LDe = np.full(ncell, 5.0e-10)
    m1, m2 = 1.0, 1836.15
    P0 = m1 * (w1 * v1).sum(axis=1) + m2 * (w2 * v2).sum(axis=1)
    K0 = m1 * (w1 * v1 ** 2).sum() + m2 * (w2 * v2 ** 2).sum()
    dp_e0 = m1 * (w1 * v1[0]).sum()
    orc.lib().orc_rng_seed(2)
    npairs = orc.coulomb_inter(cs1, v1, w1, dens1, m1, -1.0, cs2, v2, w2, dens2, m2, 1.0, LDe, cellV, 10.0, 1, False, 11,
                               40 * DT_SEC)
    assert npairs == ncell * max(n1c, n2c)
    P1 = m1 * (w1 * v1).sum(axis=1) + m2 * (w2 * v2).sum(axis=1)
    K1 = m1 * (w1 * v1 ** 2).sum() + m2 * (w2 * v2 ** 2).sum()
    exchanged = abs(m1 * (w1 * v1[0]).sum() - dp_e0)
    assert exchanged > 0.02 * abs(dp_e0)                       # the drift really slowed down
    assert abs(P1[0] - P0[0]) < 0.2 * exchanged                # ... and the ions took it, on average
    assert abs(K1 - K0) / K0 < 0.02


@pytest.mark.parametrize("angular", [3, 4])
@pytest.mark.parametrize("relativistic", [False, True])
def test_coulomb_inter_full_angle_models_conserve_for_equal_weights(angular, relativistic):
    """NANBU_FAS / NANBU_FAS_v2 through the inter-species cell driver, Galilean and LorentzScatter builds: with equal
    weights every pair conserves momentum and energy, whatever branch of the full-angle model it took."""
    rng = np.random.default_rng(23)
    ncell, n1c, n2c = 40, 24, 16
    cs1 = np.arange(ncell + 1, dtype=np.int64) * n1c
    cs2 = np.arange(ncell + 1, dtype=np.int64) * n2c
    m1, m2 = 3672.3, 5508.0                              # two ion species (the reference keeps electrons out by default)
    v1 = rng.standard_normal((3, ncell * n1c)) * 4.0e-4
    v2 = rng.standard_normal((3, ncell * n2c)) * 3.0e-4
    w1 = np.full(ncell * n1c, 2.0e27); w2 = np.full(ncell * n2c, 2.0e27)
    cellV = 1.0e-3
    dens1 = np.add.reduceat(w1, cs1[:-1]) / cellV
    dens2 = np.add.reduceat(w2, cs2[:-1]) / cellV
    LDe = np.full(ncell, 5.0e-10)
    orc.set_relativistic(relativistic)
    try:
        gam = (lambda v: np.sqrt(1.0 + (v ** 2).sum(axis=0))) if relativistic else (lambda v: 0.5 * (v ** 2).sum(axis=0))
        P0 = [m1 * v1[:, cs1[c]:cs1[c + 1]].sum(axis=1) + m2 * v2[:, cs2[c]:cs2[c + 1]].sum(axis=1) for c in range(ncell)]
        E0 = [m1 * gam(v1[:, cs1[c]:cs1[c + 1]]).sum() + m2 * gam(v2[:, cs2[c]:cs2[c + 1]]).sum() for c in range(ncell)]
        v1_0 = v1.copy()
        orc.lib().orc_rng_seed(4)
        npairs = orc.coulomb_inter(cs1, v1, w1, dens1, m1, 1.0, cs2, v2, w2, dens2, m2, 1.0, LDe, cellV, 5.0, angular, False,
                                   11, 2000 * DT_SEC)
        assert npairs == ncell * max(n1c, n2c)
        assert np.any(v1 != v1_0)
        for c in range(ncell):
            P1 = m1 * v1[:, cs1[c]:cs1[c + 1]].sum(axis=1) + m2 * v2[:, cs2[c]:cs2[c + 1]].sum(axis=1)
            E1 = m1 * gam(v1[:, cs1[c]:cs1[c + 1]]).sum() + m2 * gam(v2[:, cs2[c]:cs2[c + 1]]).sum()
            scale = m1 * np.abs(v1[:, cs1[c]:cs1[c + 1]]).sum() + m2 * np.abs(v2[:, cs2[c]:cs2[c + 1]]).sum()
            assert np.max(np.abs(P1 - P0[c])) / scale < 1e-13
            assert abs(E1 - E0[c]) / abs(E0[c]) < 1e-12
    finally:
        orc.set_relativistic(False)


def test_coulomb_enforce_conservations_intra_and_inter():
    """scattering.coulomb.enforce_conservations (Coulomb.cpp:596-714, 1182-1430): with unequal weights the weight-rejection
    update conserves momentum and energy only on average; the fix-up (shift by the weighted mean momentum change, then
    modEnergyPairwise sweeps) restores both per cell to round-off."""
    rng = np.random.default_rng(9)
    ncell, n1c, n2c = 120, 30, 18
    cs1 = np.arange(ncell + 1, dtype=np.int64) * n1c
    cs2 = np.arange(ncell + 1, dtype=np.int64) * n2c
    cellV = 1.0e-3
    LDe = np.full(ncell, 5.0e-10)
    m1, m2 = 1.0, 1836.15

    def make():
        r = np.random.default_rng(10)
        v1 = r.standard_normal((3, ncell * n1c)) * 0.02
        v2 = r.standard_normal((3, ncell * n2c)) * 0.0008
        w1 = np.where(r.random(ncell * n1c) < 0.5, 1.0e27, 3.0e27)
        w2 = np.where(r.random(ncell * n2c) < 0.5, 2.0e27, 0.5e27)
        return v1, v2, w1, w2

    def cell_P_K(v, w, cs, m):
        P = np.stack([np.add.reduceat(m * w * v[q], cs[:-1]) for q in range(3)])
        K = np.add.reduceat(m * w * (v ** 2).sum(0), cs[:-1])
        return P, K

    for on in (False, True):
        v1, v2, w1, w2 = make()
        dens1 = np.add.reduceat(w1, cs1[:-1]) / cellV
        dens2 = np.add.reduceat(w2, cs2[:-1]) / cellV
        Pa0, Ka0 = cell_P_K(v1, w1, cs1, m1)
        orc.lib().orc_rng_seed(4)
        orc.coulomb_set_enforce(on)
        try:
            orc.coulomb_intra(cs1, v1, w1, dens1, LDe, cellV, m1, -1.0, 10.0, 1, False, 11, 40 * DT_SEC)
            Pa1, Ka1 = cell_P_K(v1, w1, cs1, m1)
            Pb0 = Pa1 + cell_P_K(v2, w2, cs2, m2)[0]
            Kb0 = Ka1 + cell_P_K(v2, w2, cs2, m2)[1]
            orc.coulomb_inter(cs1, v1, w1, dens1, m1, -1.0, cs2, v2, w2, dens2, m2, 1.0, LDe, cellV, 10.0, 1, False, 11,
                              40 * DT_SEC)
            Pb1 = cell_P_K(v1, w1, cs1, m1)[0] + cell_P_K(v2, w2, cs2, m2)[0]
            Kb1 = cell_P_K(v1, w1, cs1, m1)[1] + cell_P_K(v2, w2, cs2, m2)[1]
        finally:
            orc.coulomb_set_enforce(False)
        scaleP = m1 * 3.0e27 * 0.02 * n1c
        errs = (np.abs(Pa1 - Pa0).max() / scaleP, np.abs(Ka1 - Ka0).max() / Ka0.max(),
                np.abs(Pb1 - Pb0).max() / scaleP, np.abs(Kb1 - Kb0).max() / Kb0.max())
        if on:
            assert max(errs) < 1e-12, errs
        else:
            assert min(errs) > 1e-6, errs      # without the fix-up every cell drifts


def test_coulomb_sk08_conserves_weighted_energy_per_cell():
    """Coulomb weight_method = CONSERVATIVE (Sentoku-Kemp 2008, Coulomb.cpp:730-917, 1439-1640): unequal weights, every
    pair keeps its weighted energy exactly (the heavier particle's transverse kick, Coulomb.H:796-823) and its weighted
    momentum on average."""
    ncell, n1c, n2c = 150, 30, 18
    cs1 = np.arange(ncell + 1, dtype=np.int64) * n1c
    cs2 = np.arange(ncell + 1, dtype=np.int64) * n2c
    cellV, LDe, m1, m2 = 1.0e-3, np.full(ncell, 5.0e-10), 1.0, 1836.15
    r = np.random.default_rng(12)
    v1 = r.standard_normal((3, ncell * n1c)) * 0.02
    v2 = r.standard_normal((3, ncell * n2c)) * 0.0008
    v1[0] += 0.01
    w1 = np.where(r.random(ncell * n1c) < 0.5, 1.0e27, 3.0e27)
    w2 = np.where(r.random(ncell * n2c) < 0.5, 2.0e27, 0.5e27)
    dens1 = np.add.reduceat(w1, cs1[:-1]) / cellV
    dens2 = np.add.reduceat(w2, cs2[:-1]) / cellV
    K = lambda v, w, cs, m: np.add.reduceat(m * w * (v ** 2).sum(0), cs[:-1])
    a0 = v1.copy()
    K0 = K(v1, w1, cs1, m1)
    orc.lib().orc_rng_seed(6)
    orc.coulomb_set_weight_method(True)
    try:
        orc.coulomb_intra(cs1, v1, w1, dens1, LDe, cellV, m1, -1.0, 10.0, 1, False, 11, 40 * DT_SEC)
        K1 = K(v1, w1, cs1, m1)
        Kb0 = K1 + K(v2, w2, cs2, m2)
        P0 = m1 * (w1 * v1).sum(1) + m2 * (w2 * v2).sum(1)
        d0 = m1 * (w1 * v1[0]).sum()
        orc.coulomb_inter(cs1, v1, w1, dens1, m1, -1.0, cs2, v2, w2, dens2, m2, 1.0, LDe, cellV, 10.0, 1, False, 11,
                          40 * DT_SEC)
    finally:
        orc.coulomb_set_weight_method(False)
    Kb1 = K(v1, w1, cs1, m1) + K(v2, w2, cs2, m2)
    assert np.abs(K1 - K0).max() < 1e-13 * K0.max() and np.abs(Kb1 - Kb0).max() < 1e-13 * Kb0.max()
    assert np.mean(np.any(v1 != a0, axis=0)) > 0.9
    P1 = m1 * (w1 * v1).sum(1) + m2 * (w2 * v2).sum(1)
    exchanged = abs(m1 * (w1 * v1[0]).sum() - d0)
    assert exchanged > 0.02 * abs(d0) and abs(P1[0] - P0[0]) < 0.2 * exchanged


def test_elastic_sigma_lookup():
    E = np.array([0.01, 0.1, 1.0, 10.0, 100.0])
    Q = np.array([1.0e-19, 2.0e-19, 5.0e-20, 2.0e-20, 1.0e-20])
    XI = np.array([0.0, 0.1, 0.3, 0.6, 0.9])
    mu = 1.0 * 7294.3 / (1.0 + 7294.3)
    mcSq = 9.10938370e-31 * 2.99792458e+08 ** 2 / 1.60217663e-19
    beta = lambda KE: np.sqrt(2.0 * KE / (mu * mcSq))
    # constant cross section
    assert orc.elastic_sigma(0.01, mu, const_sigma=3e-20) == (3e-20, 0.0)
    # linear interpolation inside the table, at the exact form of MathUtils::linearInterp
    s, xi = orc.elastic_sigma(beta(0.55), mu, E=E, Q=Q, XI=XI, angular=0)
    assert abs(s - (Q[2] * (0.55 - 0.1) + Q[1] * (1.0 - 0.55)) / 0.9) < 1e-30 and xi == 0.0
    # above the table: SigM ~ ln(E)/E (ISOTROPIC), SigT ~ 1/E with the last xi (OKHRIMOVSKYY)
    s, _ = orc.elastic_sigma(beta(400.0), mu, E=E, Q=Q, XI=XI, angular=0)
    assert abs(s - Q[-1] * np.log(400.0) / np.log(100.0) * 100.0 / 400.0) < 1e-32
    s, xi = orc.elastic_sigma(beta(400.0), mu, E=E, Q=Q, XI=XI, angular=1)
    assert abs(s - Q[-1] * 100.0 / 400.0) < 1e-32 and xi == XI[-1]
    # the reference's log interpolations weight the FAR node (ScatteringUtils.cpp:104-145): kept as is
    s, xi = orc.elastic_sigma(beta(0.1 * 1.0000001), mu, E=E, Q=Q, XI=XI, angular=1, loglog=True)
    assert abs(s - Q[2]) / Q[2] < 1e-5 and abs(xi - XI[2]) < 1e-5


def test_elastic_probability_and_conservation():
    rng = np.random.default_rng(4)
    ncell, n1c, n2c = 300, 20, 10
    cs1 = np.arange(ncell + 1, dtype=np.int64) * n1c
    cs2 = np.arange(ncell + 1, dtype=np.int64) * n2c
    v1 = rng.standard_normal((3, ncell * n1c)) * 0.02
    v2 = rng.standard_normal((3, ncell * n2c)) * 0.0002
    w1 = np.full(ncell * n1c, 1.0e20)
    w2 = np.full(ncell * n2c, 1.0e20)
    dens2 = np.full(ncell, 1.0e22)
    m1, m2 = 1.0, 7294.3
    sigma = 1.0e-19
    v10, v20 = v1.copy(), v2.copy()
    orc.lib().orc_rng_seed(9)
    dt = 2.0e-12
    ncoll = orc.elastic(cs1, v1, w1, m1, cs2, v2, w2, dens2, m2, dt, const_sigma=sigma)
    g = np.linalg.norm(v10, axis=0)          # partners are ~at rest
    expect = (1.0 - np.exp(-g * 2.99792458e8 * sigma * 1.0e22 * dt)).sum()
    assert abs(ncoll - expect) < 4.0 * np.sqrt(expect)
    P0 = m1 * v10.sum(axis=1) + m2 * v20.sum(axis=1)
    P1 = m1 * v1.sum(axis=1) + m2 * v2.sum(axis=1)
    assert np.max(np.abs(P1 - P0)) < 1e-12 * (m1 * np.abs(v10).sum())
    K0 = m1 * (v10 ** 2).sum() + m2 * (v20 ** 2).sum()
    K1 = m1 * (v1 ** 2).sum() + m2 * (v2 ** 2).sum()
    assert abs(K1 - K0) / K0 < 1e-12


# ---- RELATIVISTIC_PARTICLES build of TakizukaAbe: LorentzScatter (TakizukaAbe.cpp:580-659) ---------------------
def _lorentz(up1, up2, m1, m2, den, dt, b90, clog, g, ut, up):
    import ctypes as C
    f = orc.lib().orc_ta_lorentz_scatter
    f.argtypes = [C.c_void_p, C.c_void_p] + [C.c_double] * 9
    a, b = np.ascontiguousarray(up1, dtype=np.float64).copy(), np.ascontiguousarray(up2, dtype=np.float64).copy()
    small = f(orc._ptr(a), orc._ptr(b), m1, m2, den, dt, b90, clog, g, ut, up)
    return a, b, small


def test_lorentz_scatter_conserves_four_momentum_and_has_the_galilean_limit():
    import ctypes as C
    lib = orc.lib()
    lib.orc_ta_b90_fact_rel.restype = C.c_double
    lib.orc_ta_b90_fact_rel.argtypes = [C.c_double, C.c_double]
    lib.orc_ta_b90_fact.argtypes = [C.c_double] * 4
    rng = np.random.default_rng(91)
    m1, m2 = 1.0, 1836.15
    b90r = lib.orc_ta_b90_fact_rel(-1.0, 1.0)
    for scale in (2.0, 0.3):
        for _ in range(200):
            u1, u2 = rng.standard_normal(3) * scale, rng.standard_normal(3) * scale * 0.05
            a, b, _ = _lorentz(u1, u2, m1, m2, 1e30, 1e-17, b90r, 10.0, rng.standard_normal(), rng.random(), rng.random())
            g = lambda u: np.sqrt(1.0 + (u ** 2).sum())
            p0, p1 = m1 * u1 + m2 * u2, m1 * a + m2 * b
            e0, e1 = m1 * g(u1) + m2 * g(u2), m1 * g(a) + m2 * g(b)
            assert np.max(np.abs(p1 - p0)) <= 2e-13 * max(np.max(np.abs(p0)), m2 * 1e-3)
            assert abs(e1 - e0) <= 1e-14 * e0
    # |u| << 1: the same kick as computeDeltaU + the mu/m updates of the default build, same random numbers
    b90 = lib.orc_ta_b90_fact(-1.0, 1.0, m1, m2)
    mu = m1 * m2 / (m1 + m2)
    for _ in range(50):
        u1, u2 = rng.standard_normal(3) * 1e-3, rng.standard_normal(3) * 1e-5
        gss, ut, up = rng.standard_normal(), rng.random(), rng.random()
        a, b, small = _lorentz(u1, u2, m1, m2, 1e19, 1e-16, b90r, 10.0, gss, ut, up)
        dU = np.zeros(3)
        lib.orc_ta_delta_u.argtypes = [C.c_void_p, C.c_double, C.c_void_p] + [C.c_double] * 7 + [C.c_void_p]
        lib.orc_ta_delta_u(orc._ptr(u1), 1e19, orc._ptr(u2), 1e19, b90, 10.0, 1e-16, gss, ut, up, orc._ptr(dU))
        assert small == 1
        assert np.max(np.abs(a - (u1 + mu / m1 * dU))) <= 3e-6 * np.max(np.abs(dU)) + 1e-18


def test_coulomb_lorentz_scatter_conserves_and_has_the_galilean_limit():
    """Coulomb::LorentzScatter (Coulomb.cpp:1694-1793): total momentum exactly and total energy to round-off when
    both particles scatter; particle 2 untouched when the weight rejection says so; for slow particles the update
    is GalileanScatter's (mu/m1) deltaU with the same draws."""
    rng = np.random.default_rng(9)
    m1, m2 = 1.0, 1836.15
    for angular in (0, 1, 2, 5):
        for _ in range(50):
            up1 = rng.standard_normal(3) * 0.8
            up2 = rng.standard_normal(3) * 0.02
            args = (-1.0, 1.0, m1, m2, 0.0, 10.0, angular, 1.0e28, 1.0e-8, 1.0e-18, 1.0e-16)
            draws = (rng.standard_normal(), rng.random(), rng.random())
            a, b, live, s12 = orc.coulomb_lorentz_scatter(up1, up2, 1, *args, *draws)
            assert live == 1 and s12 > 0
            p0 = m1 * up1 + m2 * up2
            e0 = m1 * np.sqrt(1 + up1 @ up1) + m2 * np.sqrt(1 + up2 @ up2)
            assert np.allclose(m1 * a + m2 * b, p0, rtol=0, atol=1e-13 * np.abs(p0).max())
            e1 = m1 * np.sqrt(1 + a @ a) + m2 * np.sqrt(1 + b @ b)
            assert abs(e1 - e0) < 1e-12 * e0
            assert np.linalg.norm(a - up1) > 0
            a2, b2, _, _ = orc.coulomb_lorentz_scatter(up1, up2, 0, *args, *draws)
            assert np.array_equal(a2, a) and np.array_equal(b2, up2)
    # Galilean limit
    for angular in (0, 1, 5):
        up1 = rng.standard_normal(3) * 1e-4
        up2 = rng.standard_normal(3) * 1e-5
        args = (-1.0, 1.0, m1, m2, 0.0, 10.0, angular, 1.0e20, 1.0e-8, 1.0e-18, 1.0e-16)
        draws = (0.7, 0.3, 0.6)
        a, b, live, s12 = orc.coulomb_lorentz_scatter(up1, up2, 1, *args, *draws)
        dU, s12g = orc.coulomb_delta_u(up1, up2, *args, *draws)
        mu = m1 * m2 / (m1 + m2)
        assert abs(s12 - s12g) < 1e-6 * s12g
        assert np.allclose(a - up1, mu / m1 * dU, rtol=0, atol=1e-6 * np.linalg.norm(dU))
        assert np.allclose(b - up2, -mu / m2 * dU, rtol=0, atol=1e-6 * np.linalg.norm(dU) * mu / m2 + 1e-18)
    # identical velocities: early return
    _, _, live, _ = orc.coulomb_lorentz_scatter([0.1, 0, 0], [0.1, 0, 0], 1, -1.0, -1.0, 1.0, 1.0, 0.0, 10.0, 0, 1e28,
                                                1e-8, 1e-18, 1e-16, 0.1, 0.2, 0.3)
    assert live == 0


def test_coulomb_intra_relativistic_build_conserves_for_equal_weights():
    """applyIntraScattering_PROB of the RELATIVISTIC_PARTICLES build: every pair goes through LorentzScatter; with
    equal weights both partners scatter, so the cell's momentum and energy (sum m gamma) are conserved."""
    rng = np.random.default_rng(10)
    ncell, npc = 6, 9
    n = ncell * npc
    cs = np.arange(0, n + 1, npc)
    v = rng.standard_normal((3, n)) * 0.5
    w = np.full(n, 2.0e10)
    dens = np.full(ncell, 1.0e28)
    LDe = np.full(ncell, 1.0e-8)
    orc.set_relativistic(True)
    try:
        orc.lib().orc_rng_seed(5)
        v1 = v.copy()
        orc.coulomb_intra(cs, v1, w, dens, LDe, 1.0e-18, 1.0, -1.0, 10.0, 1, False, 11, 1.0e-16)
    finally:
        orc.set_relativistic(False)
    assert np.max(np.abs(v1 - v)) > 1e-6
    for c in range(ncell):
        sl = slice(cs[c], cs[c + 1])
        assert np.allclose(v1[:, sl].sum(1), v[:, sl].sum(1), rtol=0, atol=1e-12)
        g0 = np.sqrt(1 + (v[:, sl] ** 2).sum(0)).sum()
        g1 = np.sqrt(1 + (v1[:, sl] ** 2).sum(0)).sum()
        assert abs(g1 - g0) < 1e-12 * g0


def _hs_cells(rng, ncell, npc, vth, mass, wgt, Vc):
    n = ncell * npc
    cs = np.arange(0, n + 1, npc)
    v = rng.standard_normal((3, n)) * vth
    w = np.full(n, wgt)
    dens = np.full(ncell, npc * wgt / Vc)
    ene = np.zeros((3, ncell))
    for c in range(ncell):
        ene[:, c] = 0.5 * mass * (w[cs[c]:cs[c + 1]] * v[:, cs[c]:cs[c + 1]] ** 2).sum(1) / Vc
    return cs, v, w, dens, ene


def test_hard_sphere_self_conserves_and_accepts_like_a_maxwellian():
    """HardSphere::applySelfScattering (HardSphere.cpp:223-418): equal weights -> every accepted pair conserves
    momentum and energy; acceptance = <g>/gmax = (4/sqrt(pi)) vth / (5 vth) ~ 0.45 for a Maxwellian."""
    rng = np.random.default_rng(21)
    ncell, npc, mass, Vc = 200, 40, 1836.0, 1.0e-9
    cs, v, w, dens, ene = _hs_cells(rng, ncell, npc, 1.0e-3, mass, 1.0e20, Vc)
    sig = orc.hs_sigmaT(1.0e-10, 1.0e-10)
    assert abs(sig - np.pi * 4.0e-20) < 1e-30
    gmax = 5.0 * 1.0e-3 * 2.99792458e8
    dt = 0.8 / (dens[0] * sig * gmax)                      # nuMax*dt ~ 0.8
    v0 = v.copy()
    orc.lib().orc_rng_seed(3)
    ncand, ncoll = orc.hs_self(cs, v, w, dens, ene, mass, sig, dt)
    assert abs(ncand - ncell * 0.5 * (npc - 1) * 0.8) < 0.05 * ncand
    assert 0.40 < ncoll / ncand < 0.50
    dv = v.reshape(3, ncell, npc).sum(2) - v0.reshape(3, ncell, npc).sum(2)
    assert np.max(np.abs(dv)) < 1e-16
    e0, e1 = (v0 ** 2).reshape(3, ncell, npc).sum((0, 2)), (v ** 2).reshape(3, ncell, npc).sum((0, 2))
    assert np.max(np.abs(e1 - e0) / e0) < 1e-13
    assert np.mean(np.any(v != v0, axis=0)) > 0.2


def test_hard_sphere_inter_conserves_total_momentum_and_energy():
    rng = np.random.default_rng(22)
    ncell, Vc = 120, 1.0e-9
    m1, m2 = 4.0 * 1836.0, 40.0 * 1836.0
    cs1, v1, w1, d1, e1 = _hs_cells(rng, ncell, 30, 2.0e-3, m1, 1.0e20, Vc)
    cs2, v2, w2, d2, e2 = _hs_cells(rng, ncell, 20, 5.0e-4, m2, 1.0e20, Vc)
    sig = orc.hs_sigmaT(1.2e-10, 1.8e-10)
    a1, a2 = v1.copy(), v2.copy()
    orc.lib().orc_rng_seed(4)
    ncand, ncoll = orc.hs_inter(cs1, v1, w1, d1, e1, m1, cs2, v2, w2, d2, e2, m2, Vc, sig, 3.0e-18)
    assert 1000 < ncand < 1000000 and 0.2 < ncoll / ncand < 0.7
    for c in range(0, ncell, 7):
        s1, s2 = slice(cs1[c], cs1[c + 1]), slice(cs2[c], cs2[c + 1])
        p0 = m1 * a1[:, s1].sum(1) + m2 * a2[:, s2].sum(1)
        p1 = m1 * v1[:, s1].sum(1) + m2 * v2[:, s2].sum(1)
        assert np.max(np.abs(p1 - p0)) < 1e-12 * np.max(np.abs(p0)) + 1e-14
        k0 = m1 * (a1[:, s1] ** 2).sum() + m2 * (a2[:, s2] ** 2).sum()
        k1 = m1 * (v1[:, s1] ** 2).sum() + m2 * (v2[:, s2] ** 2).sum()
        assert abs(k1 - k0) < 1e-12 * k0


def test_vhs_self_conserves_and_reduces_to_hard_sphere_for_eta_one_half_limit():
    """VariableHardSphere::applySelfScattering (VariableHardSphere.cpp:217-412): both partners scatter, so every
    cell conserves momentum and energy; acceptance = <g sigma(g)> / (gmax sigma(gmax)) in (0, 1)."""
    rng = np.random.default_rng(23)
    ncell, npc, Vc = 150, 40, 1.0e-9
    mass = 39.948 * 1822.888
    cs, v, w, dens, ene = _hs_cells(rng, ncell, npc, 5.0e-6, mass, 1.0e20, Vc)
    fourPiA, fourOverAlpha = orc.vhs_consts(mass, 0.81, 273.0, 2.117e-5)
    assert abs(fourOverAlpha - (2 * 0.81 - 1)) < 1e-15
    gmax = 5.0 * 5.0e-6 * 2.99792458e8
    sigmax = fourPiA * gmax ** (-fourOverAlpha)
    assert 1e-20 < sigmax < 1e-17                          # a molecular cross section [m^2]
    dt = 0.7 / (dens[0] * sigmax * gmax)
    v0 = v.copy()
    orc.lib().orc_rng_seed(6)
    ncand, ncoll = orc.vhs_self(cs, v, dens, ene, mass, fourPiA, fourOverAlpha, dt)
    assert abs(ncand - ncell * 0.5 * (npc - 1) * 0.7) < 0.06 * ncand
    assert 0.3 < ncoll / ncand < 0.9
    dv = v.reshape(3, ncell, npc).sum(2) - v0.reshape(3, ncell, npc).sum(2)
    assert np.max(np.abs(dv)) < 1e-18
    e0, e1 = (v0 ** 2).reshape(3, ncell, npc).sum((0, 2)), (v ** 2).reshape(3, ncell, npc).sum((0, 2))
    assert np.max(np.abs(e1 - e0) / e0) < 1e-13


def test_hard_sphere_conservative_weight_method_keeps_weight_momentum_and_energy():
    """HardSphere::applySelfScattering with weight_method = CONSERVATIVE (HardSphere.cpp:357-392): unequal-weight
    pairs go through ScatteringUtils::collapseThreeToTwo (pinned on the reference in test_ref_pin.py); every cell keeps
    its total weight, its weighted momentum and its weighted energy to round-off -- which the PROBABILISTIC method
    only does on average."""
    rng = np.random.default_rng(31)
    ncell, npc, mass, Vc = 100, 30, 1836.0, 1.0e-9
    cs, v, w, dens, ene = _hs_cells(rng, ncell, npc, 1.0e-3, mass, 1.0e20, Vc)
    w *= rng.choice([0.5, 1.0, 2.0], size=w.size)
    for c in range(ncell):
        sl = slice(cs[c], cs[c + 1])
        dens[c] = w[sl].sum() / Vc
        ene[:, c] = 0.5 * mass * (w[sl] * v[:, sl] ** 2).sum(1) / Vc
    sig = orc.hs_sigmaT(1.0e-10, 1.0e-10)
    dt = 0.8 / (dens.max() * sig * 5.0e-3 * 2.99792458e8)
    v0, w0 = v.copy(), w.copy()
    orc.lib().orc_rng_seed(3)
    ncand, ncoll = orc.hs_self_conservative(cs, v, w, dens, ene, mass, sig, dt)
    assert ncoll > 200 and not np.isnan(v).any()
    assert np.mean(w != w0) > 0.05
    for c in range(ncell):
        sl = slice(cs[c], cs[c + 1])
        assert abs(w[sl].sum() - w0[sl].sum()) < 1e-14 * w0[sl].sum()
        p0, p1 = (w0[sl] * v0[:, sl]).sum(1), (w[sl] * v[:, sl]).sum(1)
        assert np.max(np.abs(p1 - p0)) < 1e-13 * np.abs(w0[sl] * v0[:, sl]).sum(1).max()
        e0, e1 = (w0[sl] * v0[:, sl] ** 2).sum(), (w[sl] * v[:, sl] ** 2).sum()
        assert abs(e1 - e0) < 1e-13 * e0
    # the PROBABILISTIC method on the same cells does not conserve per cell
    v2, w2 = v0.copy(), w0.copy()
    orc.lib().orc_rng_seed(3)
    orc.hs_self(cs, v2, w2, dens, ene, mass, sig, dt)
    drift = max(abs((w0[cs[c]:cs[c + 1]] * v2[:, cs[c]:cs[c + 1]] ** 2).sum() - (w0[cs[c]:cs[c + 1]] * v0[:, cs[c]:cs[c + 1]] ** 2).sum())
                / (w0[cs[c]:cs[c + 1]] * v0[:, cs[c]:cs[c + 1]] ** 2).sum() for c in range(ncell))
    assert drift > 1e-6


def test_elastic_conservative_weight_method_keeps_weight_momentum_and_energy():
    """Elastic::electronImpact with weight_method = CONSERVATIVE (Elastic.cpp:333-358), oracle only so far: light
    projectiles (electrons, small weights) on heavier-weight targets; every cell keeps the targets' total weight and the
    pair's weighted momentum m1 w1 v1 + m2 w2 v2 and weighted energy to round-off."""
    rng = np.random.default_rng(41)
    ncell, Vc = 60, 1.0e-9
    m1, m2 = 1.0, 40.0 * 1836.0
    cs1, v1, w1, d1, _ = _hs_cells(rng, ncell, 25, 2.0e-2, m1, 0.25e20, Vc)
    cs2, v2, w2, d2, _ = _hs_cells(rng, ncell, 15, 2.0e-5, m2, 1.0e20, Vc)
    w2 *= rng.choice([1.0, 2.0], size=w2.size)             # the targets' weights differ, so the merges change them
    d2 = np.array([w2[cs2[c]:cs2[c + 1]].sum() / Vc for c in range(ncell)])
    a1, a2, u2 = v1.copy(), v2.copy(), w2.copy()
    orc.lib().orc_rng_seed(8)
    ncoll = orc.elastic_conservative(cs1, v1, w1, m1, cs2, v2, w2, d2, m2, 2.0e-11, 1.0e-19)
    assert ncoll > 100 and not np.isnan(v2).any()
    assert np.mean(w2 != u2) > 0.02
    for c in range(ncell):
        s1, s2 = slice(cs1[c], cs1[c + 1]), slice(cs2[c], cs2[c + 1])
        assert abs(w2[s2].sum() - u2[s2].sum()) < 1e-14 * u2[s2].sum()
        p0 = m1 * (w1[s1] * a1[:, s1]).sum(1) + m2 * (u2[s2] * a2[:, s2]).sum(1)
        p1 = m1 * (w1[s1] * v1[:, s1]).sum(1) + m2 * (w2[s2] * v2[:, s2]).sum(1)
        scale = m1 * np.abs(w1[s1] * a1[:, s1]).sum() + m2 * np.abs(u2[s2] * a2[:, s2]).sum()
        assert np.max(np.abs(p1 - p0)) < 1e-13 * scale
        e0 = m1 * (w1[s1] * a1[:, s1] ** 2).sum() + m2 * (u2[s2] * a2[:, s2] ** 2).sum()
        e1 = m1 * (w1[s1] * v1[:, s1] ** 2).sum() + m2 * (w2[s2] * v2[:, s2] ** 2).sum()
        assert abs(e1 - e0) < 1e-12 * e0


def test_hard_sphere_inter_conservative_weight_method():
    """HardSphere between species with weight_method CONSERVATIVE (HardSphere.cpp:594-636): every cell keeps the total weight
    of either species; with EQUAL masses (where the reference's 0.5 deltaU equals mu/m deltaU) it also keeps the weighted
    momentum and the weighted energy of the two species together, per direction-summed, to round-off."""
    rng = np.random.default_rng(17)
    ncell = 40
    c1 = rng.choice([0, 1, 2, 9, 30], size=ncell)
    c2 = rng.choice([0, 1, 3, 12, 25], size=ncell)
    cs1 = np.concatenate([[0], np.cumsum(c1)]).astype(np.int64)
    cs2 = np.concatenate([[0], np.cumsum(c2)]).astype(np.int64)
    n1, n2 = int(cs1[-1]), int(cs2[-1])
    v1, v2 = rng.standard_normal((3, n1)) * 2e-3, rng.standard_normal((3, n2)) * 1e-3
    w1 = 1.0e27 * rng.choice([0.5, 1.0], size=n1)
    w2 = 1.0e27 * rng.choice([1.0, 2.0], size=n2)
    Vc = 1.0e-3
    m = 4.0 * 1836.15

    def moments(cs, v, w, mass):
        dens = np.array([w[cs[c]:cs[c + 1]].sum() / Vc for c in range(ncell)])
        ene = np.array([[0.5 * mass * (w[cs[c]:cs[c + 1]] * v[d, cs[c]:cs[c + 1]] ** 2).sum() / Vc for c in range(ncell)]
                        for d in range(3)])
        return dens, ene
    d1, e1 = moments(cs1, v1, w1, m)
    d2, e2 = moments(cs2, v2, w2, m)
    sig = orc.hs_sigmaT(1.2e-10, 1.8e-10)
    a1, a2, b1, b2 = v1.copy(), v2.copy(), w1.copy(), w2.copy()
    orc.lib().orc_rng_seed(9)
    ncand, ncoll = orc.hs_inter_conservative(cs1, a1, b1, d1, e1, m, cs2, a2, b2, d2, e2, m, Vc, sig, 3.0e-15)
    assert ncoll > 100 and not np.isnan(a1).any() and not np.isnan(a2).any()
    assert np.mean(b2 != w2) > 0.02 or np.mean(b1 != w1) > 0.02          # merges happened
    for c in range(ncell):
        s1, s2 = slice(cs1[c], cs1[c + 1]), slice(cs2[c], cs2[c + 1])
        assert abs(b1[s1].sum() - w1[s1].sum()) <= 1e-13 * max(w1[s1].sum(), 1.0)
        assert abs(b2[s2].sum() - w2[s2].sum()) <= 1e-13 * max(w2[s2].sum(), 1.0)
        p0 = (w1[s1] * v1[:, s1]).sum(1) + (w2[s2] * v2[:, s2]).sum(1)
        p1 = (b1[s1] * a1[:, s1]).sum(1) + (b2[s2] * a2[:, s2]).sum(1)
        k0 = (w1[s1] * v1[:, s1] ** 2).sum() + (w2[s2] * v2[:, s2] ** 2).sum()
        k1 = (b1[s1] * a1[:, s1] ** 2).sum() + (b2[s2] * a2[:, s2] ** 2).sum()
        if k0 > 0:
            assert np.max(np.abs(p1 - p0)) < 1e-11 * np.sqrt(k0 * (w1[s1].sum() + w2[s2].sum()))
            assert abs(k1 - k0) < 1e-11 * k0
