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
