"""The reference's RELATIVISTIC_PARTICLES build of the push as a per-species flag (pgpu_species_desc.relativistic,
.higuera_cary): Boris with the time-centred gamma (pinned bit for bit on the reference's own applyForces through
the oracle, tests/test_ref_pin.py), positions and the Picard step norm with getImplicitGamma, current deposit
with w/gamma, setStableDt and globalMoments.  GPU (C ABI) against the oracle in relativistic mode; velocities up
to gamma*beta ~ 1."""
import numpy as np
import pytest

from common import INTERPS, Problem, orc, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _relativistic_oracle():
    orc.set_relativistic(False)
    yield
    orc.set_relativistic(False)
    from picnic_b200 import capi
    capi.load().pgpu_set_exact_math(0)


def _prob(D, seed, n=3000, max_disp=0.9):
    if D == 1:
        prob = Problem(1, (24,), (0.25,), (0.5,), 4, n, seed=seed, max_disp=max_disp, E0=0.05, B0=0.3)
    else:
        prob = Problem(2, (12, 10), (0.25, 0.3), (0.5, -1.0), 4, n, seed=seed, max_disp=max_disp, E0=0.05, B0=0.3)
    rng = np.random.default_rng(seed + 100)
    prob.vold = np.ascontiguousarray(rng.standard_normal((3, n)) * 0.7)       # gamma*beta of order one
    prob.v = np.ascontiguousarray(prob.vold + rng.standard_normal((3, n)) * 0.05)
    return prob


def _gpu(pgpu, prob, interp, hc, **kw):
    grid = pgpu.Grid(prob.D, prob.ncell, prob.xmin, prob.dx, prob.nghost, [1] * prob.D)
    E, B = prob.fields_for_gpu()
    grid.set_fields(E, B)
    sp = pgpu.Species(grid, 1.0, -1.0, kw.pop("fnorm", -0.4), kw.pop("cvac_norm", 0.9), interp_N=1, interp_J=interp,
                      interp_E=interp, relativistic=True, higuera_cary=hc, **kw)
    sp.upload(prob.x, prob.v, prob.w, xold=prob.xold, vold=prob.vold, ids=np.arange(prob.n, dtype=np.uint64))
    return grid, sp


@pytest.mark.parametrize("exact", [1, 0])
@pytest.mark.parametrize("hc", [False, True])
@pytest.mark.parametrize("half", [True, False])
def test_boris_relativistic(pgpu, hc, half, exact):
    prob = _prob(2, 51)
    pgpu.load().pgpu_set_exact_math(exact)
    grid, sp = _gpu(pgpu, prob, INTERPS["CIC"], hc)
    sp.interpolate_fields()
    Ep, Bp = sp.particle_fields()
    sp.advance_velocities(0.2, half)
    got = sp.download()["v"]
    orc.set_relativistic(True, hc)
    want = orc.boris(prob.v, prob.vold, np.ascontiguousarray(Ep), np.ascontiguousarray(Bp), -0.4, 0.2 * 0.9, half)
    if exact:
        assert np.array_equal(got, want)          # same operation order as the (pinned) oracle
    else:
        assert rel_err(got, want) <= 1e-13
    sp.destroy(); grid.destroy()


@pytest.mark.parametrize("D", [1, 2])
def test_positions_relativistic(pgpu, D):
    prob = _prob(D, 52)
    grid, sp = _gpu(pgpu, prob, INTERPS["CIC"], False)
    orc.set_relativistic(True)
    sp.advance_positions_explicit(0.3, half=True)
    x = np.zeros_like(prob.x)
    orc.lib().orc_advance_positions_explicit(D, prob.n, orc._ptr(x), orc._ptr(prob.xold), orc._ptr(prob.v), 0.5 * 0.3 * 0.9)
    assert np.array_equal(sp.download()["x"], x)
    sp.advance_positions_implicit(0.3)
    orc.advance_positions_implicit_rel(D, x, prob.xold, prob.v, prob.vold, 0.3 * 0.9)
    assert np.array_equal(sp.download()["x"], x)
    # setStableDt and globalMoments of the relativistic build
    g = np.sqrt(1.0 + (prob.v ** 2).sum(axis=0))
    want_dt = 1.0 / max(np.max(np.abs(prob.v[d] / g) / prob.dx[d]) for d in range(D)) / 0.9    # 1/maxDtinv/cvac_norm
    assert abs(sp.stable_dt() - want_dt) <= 1e-14 * want_dt
    mom = sp.global_moments()
    want_e = np.sum(prob.w * (g ** 2 - 1.0) * 2.0 / (g + 1.0))
    assert abs(mom[4] + mom[5] + mom[6] - want_e) <= 1e-12 * want_e
    sp.destroy(); grid.destroy()


@pytest.mark.parametrize("exact", [1, 0])
@pytest.mark.parametrize("interp", ["CIC", "CC0", "CC1"])
@pytest.mark.parametrize("D", [1, 2])
def test_picard_advance_and_deposit_relativistic(pgpu, D, interp, exact):
    prob = _prob(D, 53)
    pgpu.load().pgpu_set_exact_math(exact)
    grid, sp = _gpu(pgpu, prob, INTERPS[interp], False, rtol=1e-12, iter_max=30)
    st = sp.advance_iteratively(0.5, deposit=True)
    got = sp.download()
    J = [sp.current_get(c) for c in range(3)]
    orc.set_relativistic(True)
    x, v = prob.x.copy(), prob.v.copy()
    rc, apply_its, unconv, _ = orc.advance_particles_iteratively(prob.geom, INTERPS[interp], x, prob.xold, v, prob.vold,
                                                                prob.E, prob.B, -0.4, 0.5 * 0.9, 1e-12, 30)
    assert rc == 0 and st.num_unconverged == unconv and unconv <= 3
    J0 = prob.new_J()
    assert orc.deposit_current_rel(prob.geom, INTERPS[interp], x, prob.xold, v, prob.vold, prob.w, 0.5 * 0.9, J0) == 0
    tol = 1e-14 if exact else 1e-12
    assert np.max(np.abs(got["x"] - x) / np.array(prob.dx)[:, None]) <= (1e-13 if exact else 4e-12)
    assert rel_err(got["v"], v) <= (1e-14 if exact else 1e-11)
    if exact:
        assert st.num_apply_its == apply_its
    for c in range(3):
        assert rel_err(J[c], J0[c].a * (-1.0)) <= (2e-13 if exact else 1e-11), c
    # the non-relativistic answer is measurably different at these velocities
    orc.set_relativistic(False)
    x2, v2 = prob.x.copy(), prob.v.copy()
    orc.advance_particles_iteratively(prob.geom, INTERPS[interp], x2, prob.xold, v2, prob.vold, prob.E, prob.B, -0.4,
                                      0.5 * 0.9, 1e-12, 30)
    assert rel_err(v2, v) > 1e-3
    sp.destroy(); grid.destroy()


def test_deposit_from_explicit_solver_relativistic(pgpu):
    prob = _prob(2, 54)
    grid, sp = _gpu(pgpu, prob, INTERPS["CIC"], False)
    sp.set_current_density(0.1, from_explicit=True)
    J0 = prob.new_J()
    orc.set_relativistic(True)
    assert orc.deposit_current_rel(prob.geom, INTERPS["CIC"], prob.x, prob.xold, prob.v, prob.vold, prob.w, 0.09, J0,
                                   from_explicit=True) == 0
    for c in range(3):
        assert rel_err(sp.current_get(c), J0[c].a * (-1.0)) <= 1e-12
    sp.destroy(); grid.destroy()


# ---- TakizukaAbe::LorentzScatter (TakizukaAbe.cpp:580-659) ---------------------------------------------------
def _orc_lorentz(u1, u2, m1, m2, den, dt, b90, clog, g, ut, up):
    import ctypes as C
    f = orc.lib().orc_ta_lorentz_scatter
    f.argtypes = [C.c_void_p, C.c_void_p] + [C.c_double] * 9
    a, b = np.ascontiguousarray(u1).copy(), np.ascontiguousarray(u2).copy()
    f(orc._ptr(a), orc._ptr(b), m1, m2, den, dt, b90, clog, g, ut, up)
    return a, b


def test_lorentz_scatter_matches_oracle(pgpu):
    import ctypes as C
    lib = orc.lib()
    lib.orc_ta_b90_fact_rel.restype = C.c_double
    lib.orc_ta_b90_fact_rel.argtypes = [C.c_double, C.c_double]
    rng = np.random.default_rng(61)
    n = 3000
    m1, m2 = 1.0, 1836.15
    u1 = np.ascontiguousarray(rng.standard_normal((3, n)) * 0.8)
    u2 = np.ascontiguousarray(rng.standard_normal((3, n)) * 0.02)
    u1[:, :300] *= 1e-3; u2[:, :300] *= 1e-3          # slow pairs: s12 >= 2, isotropic branch
    den = np.ascontiguousarray(10.0 ** rng.uniform(26, 31, n))
    g, ut, up = rng.standard_normal(n), rng.random(n), rng.random(n)
    b90 = lib.orc_ta_b90_fact_rel(-1.0, 1.0)
    o1, o2 = np.zeros((3, n)), np.zeros((3, n))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    pgpu.check(pgpu.load().pgpu_ta_lorentz_scatter(n, p(u1), p(u2), m1, m2, p(den), 1.77e-18, b90, 10.0, p(g), p(ut),
                                                   p(up), p(o1), p(o2)))
    worst = 0.0
    for i in range(n):
        a, b = _orc_lorentz(u1[:, i], u2[:, i], m1, m2, den[i], 1.77e-18, b90, 10.0, g[i], ut[i], up[i])
        s = max(np.linalg.norm(u1[:, i]), 1e-30)
        worst = max(worst, np.max(np.abs(o1[:, i] - a)) / s, np.max(np.abs(o2[:, i] - b)) / s)
    assert worst < 1e-12, worst        # long double scalars in the reference vs doubles on the device
    gam = lambda u: np.sqrt(1.0 + (u ** 2).sum(axis=0))
    e0, e1 = m1 * gam(u1) + m2 * gam(u2), m1 * gam(o1) + m2 * gam(o2)
    assert np.max(np.abs(e1 - e0) / e0) < 1e-13
    assert np.max(np.abs(m1 * o1 + m2 * o2 - m1 * u1 - m2 * u2)) < 1e-12 * m2


def test_collide_ta_relativistic_conserves_four_momentum(pgpu):
    """pgpu_collide_ta with relativistic species: LorentzScatter per pair -- total momentum and total energy
    sum m*gamma of every cell conserved to round-off (self: odd and even cells; between species)."""
    from picnic_b200 import decks
    rng = np.random.default_rng(62)
    ncell = 40
    deck = decks.Deck(D=1, ncell=(ncell,), dx=(0.25,), xmin=(0.0,), nghost=2)
    grid = pgpu.Grid(1, (ncell,), (0.0,), (0.25,), 2, (1,), volume_scale=deck.volume_scale)

    def cells(counts):
        return np.concatenate([(c + rng.random(k)) * 0.25 for c, k in enumerate(counts)])[None, :]

    def species(mass, charge, x, scale):
        sp = pgpu.Species(grid, mass, charge, 1.0, deck.units.cvac_norm, relativistic=True)
        n = x.shape[1]
        sp.upload(x, rng.standard_normal((3, n)) * scale, np.full(n, 1e28), ids=np.arange(n, dtype=np.uint64))
        sp.bin_particles(); sp.set_moments()
        return sp

    xe = cells(rng.choice([7, 10, 21], size=ncell))
    xi = cells(rng.choice([5, 12], size=ncell))
    se, si = species(1.0, -1.0, xe, 0.6), species(1836.15, 1.0, xi, 0.01)

    def totals(sp):
        d = sp.download()
        cell = np.floor(d["x"][0] / 0.25).astype(int)
        g = np.sqrt(1.0 + (d["v"] ** 2).sum(axis=0))
        P = np.stack([np.bincount(cell, d["v"][k], ncell) for k in range(3)])
        return P, np.bincount(cell, g, ncell), d["v"]

    Pe0, Ee0, ve0 = totals(se)
    n1 = pgpu.collide_ta(se, se, 10.0, 1.77e-18, 11, 0)
    Pe1, Ee1, ve1 = totals(se)
    assert n1 > 0 and not np.array_equal(ve0, ve1)
    assert np.max(np.abs(Pe1 - Pe0)) < 1e-12 and np.max(np.abs(Ee1 - Ee0) / Ee0) < 1e-13
    Pi0, Ei0, _ = totals(si)
    n2 = pgpu.collide_ta(se, si, 10.0, 1.77e-18, 11, 1)
    Pe2, Ee2, _ = totals(se)
    Pi2, Ei2, _ = totals(si)
    assert n2 > 0
    assert np.max(np.abs((Pe2 + 1836.15 * Pi2) - (Pe1 + 1836.15 * Pi0))) < 1e-10
    tot0, tot2 = Ee1 + 1836.15 * Ei0, Ee2 + 1836.15 * Ei2
    assert np.max(np.abs(tot2 - tot0) / tot0) < 1e-13
    se.destroy(); si.destroy(); grid.destroy()
