"""The specialised 2D CC1 fused kernel (pgpu_advance_cc1.cu: single-segment closed forms,
per-warp run-sum deposit, deferred list) against the CPU oracle and against the generic
visitor kernel, on sorted, shuffled, face-hugging and ghost-region particles."""
import numpy as np
import pytest

from common import INTERPS, Problem, make_gpu, orc, rel_err

pytestmark = pytest.mark.gpu

FN, CV, DT = -0.7, 0.9986, 0.5


def _oracle(prob, rtol=1e-12, itmax=21, charge=-1.0, vs=3.0):
    x, v = prob.x.copy(), prob.v.copy()
    rc, apply_its, unconv, its = orc.advance_particles_iteratively(
        prob.geom, orc.CC1, x, prob.xold, v, prob.vold, prob.E, prob.B, FN, DT * CV, rtol, itmax)
    assert rc == 0
    J0 = prob.new_J()
    assert orc.deposit_current(prob.geom, orc.CC1, x, prob.xold, v, prob.w, DT * CV, J0) == 0
    for c in range(3):
        orc.scale_fab(J0[c], 2, charge / vs)
    return x, v, J0, apply_its, unconv


def _run(pgpu, prob, mode, rtol=1e-12, itmax=21):
    pgpu.check(pgpu.load().pgpu_set_deposit_mode(mode))
    grid, sp = make_gpu(pgpu, prob, INTERPS["CC1"], rtol=rtol, iter_max=itmax, fnorm=FN, cvac_norm=CV,
                        charge=-1.0, volume_scale=3.0)
    pgpu.profile_reset(); pgpu.profile_enable(True)
    st = sp.advance_iteratively(DT, deposit=True)
    pgpu.profile_enable(False)
    got = sp.download()
    J = [sp.current_get(c) for c in range(3)]
    launched_fast = pgpu.profile_query("advance_cc1_fused")[1]
    sp.destroy(); grid.destroy()
    pgpu.check(pgpu.load().pgpu_set_deposit_mode(1))
    return got, J, st, launched_fast


def _check(prob, got, J, x, v, J0, tol_x=4e-12, tol_v=1e-11, tol_j=1e-11):
    assert np.max(np.abs(got["x"] - x) / np.array(prob.dx)[:, None]) <= tol_x
    assert rel_err(got["v"], v) <= tol_v
    for c in range(3):
        assert rel_err(J[c], J0[c].a) <= tol_j, c


@pytest.mark.parametrize("order", ["sorted", "shuffled"])
@pytest.mark.parametrize("max_disp", [0.02, 0.5, 1.6])
def test_fast_kernel_matches_oracle(pgpu, order, max_disp):
    """max_disp 0.02: ~all particles single-segment (fast path); 0.5: a mix; 1.6: most deferred."""
    prob = Problem(2, (24, 20), (0.25, 0.3), (0.5, -1.0), 4, 20000, seed=21, max_disp=max_disp, E0=0.3, B0=0.8)
    if order == "sorted":
        cells = orc.bin_cells(prob.geom, prob.xold)
        perm = np.argsort(cells[0] + cells[1] * prob.ncell[0], kind="stable")
        for name in ("x", "xold", "v", "vold"):
            setattr(prob, name, np.ascontiguousarray(getattr(prob, name)[:, perm]))
        prob.w = np.ascontiguousarray(prob.w[perm])
    x, v, J0, apply_its, unconv = _oracle(prob)
    got, J, st, nfast = _run(pgpu, prob, 1)
    assert nfast == 1                                  # the specialised kernel did run
    _check(prob, got, J, x, v, J0)
    assert st.num_parts_its == prob.n
    assert abs(st.num_apply_its - apply_its) <= max(3, prob.n // 100)
    assert st.num_unconverged == unconv
    # and the generic kernel alone (mode 0) agrees with it to round-off
    got0, J00, st0, nfast0 = _run(pgpu, prob, 0)
    assert nfast0 == 0
    assert np.max(np.abs(got["x"] - got0["x"]) / np.array(prob.dx)[:, None]) <= 4e-12
    for c in range(3):
        assert rel_err(J[c], J00[c]) <= 1e-12


@pytest.mark.parametrize("passes", [1, 3, 6])
def test_fast_kernel_fixed_pass_count_at_1e12(pgpu, passes):
    """The tolerances above (4e-12 dx, 1e-11) are the error budget of an rtol-limited loop: a particle whose last
    step norm lands within round-off of rtol = 1e-12 stops one pass earlier in one implementation than in the other,
    and the two answers then differ by ~rtol.  With the pass count FIXED (rtol = 0: no particle ever converges, every
    particle does iter_max + 1 field applications in both) only the arithmetic differs, and the north_star bar of
    1e-12 applies -- with room to spare."""
    prob = Problem(2, (24, 20), (0.25, 0.3), (0.5, -1.0), 4, 20000, seed=31, max_disp=0.3, E0=0.3, B0=0.8)
    x, v, J0, apply_its, unconv = _oracle(prob, rtol=0.0, itmax=passes)
    got, J, st, nfast = _run(pgpu, prob, 1, rtol=0.0, itmax=passes)
    assert nfast == 1
    _check(prob, got, J, x, v, J0, tol_x=1e-12, tol_v=1e-12, tol_j=1e-12)
    assert st.num_apply_its == apply_its
    assert st.num_unconverged == unconv == prob.n


def test_fast_kernel_faces_and_ghost_region(pgpu):
    """x_old exactly on dual-cell faces (cell centres), on primal faces, and up to ghosts-1 cells
    outside the box: the same-cell decision must be the reference's, and edge stencils defer."""
    ncell, dx, xmin, ng = (16, 12), (0.25, 0.3), (0.5, -1.0), 4
    n = 6000
    prob = Problem(2, ncell, dx, xmin, ng, n, seed=22, max_disp=0.05, E0=0.3, B0=0.8)
    rng = np.random.default_rng(23)
    for d in range(2):
        k = rng.integers(-2 * (ng - 2), 2 * (ncell[d] + ng - 2), size=n)
        on_face = rng.random(n) < 0.5
        xf = xmin[d] + k * (0.5 * dx[d])                       # primal faces and cell centres
        prob.xold[d] = np.where(on_face, xf, prob.xold[d])
        eps = rng.choice([-2, -1, 0, 0, 1, 2], size=n)
        nudged = xf.copy()
        for s in (-2, -1, 1, 2):
            m = on_face & (eps == s)
            t = xf[m]
            for _ in range(abs(s)):
                t = np.nextafter(t, np.inf if s > 0 else -np.inf)
            nudged[m] = t
        prob.xold[d] = np.where(on_face, nudged, prob.xold[d])
        prob.x[d] = prob.xold[d] + (rng.random(n) * 2 - 1) * 0.02 * dx[d]
        still = rng.random(n) < 0.1                             # zero displacement
        prob.x[d] = np.where(still, prob.xold[d], prob.x[d])
    prob.vold[:2] *= 0.05                                       # keep the converged orbit short
    prob.v[:] = prob.vold
    x, v, J0, apply_its, unconv = _oracle(prob)
    got, J, st, nfast = _run(pgpu, prob, 1)
    assert nfast == 1
    _check(prob, got, J, x, v, J0)


def test_fast_kernel_single_pass_and_no_deposit(pgpu):
    prob = Problem(2, (24, 20), (0.25, 0.3), (0.5, -1.0), 4, 8000, seed=24, max_disp=0.1, E0=0.3, B0=0.8)
    grid, sp = make_gpu(pgpu, prob, INTERPS["CC1"], iter_max=0, order_swap=0, fnorm=FN, cvac_norm=CV)
    pgpu.profile_reset(); pgpu.profile_enable(True)
    sp.advance_iteratively(DT, deposit=False)        # iter_max == 0 -> advanceParticles
    pgpu.profile_enable(False)
    assert pgpu.profile_query("advance_cc1")[1] == 1
    got = sp.download()
    x, v = prob.x.copy(), prob.v.copy()
    assert orc.advance_particles(prob.geom, orc.CC1, x, prob.xold, v, prob.vold, prob.E, prob.B, FN, DT * CV, 0) == 0
    assert rel_err(got["v"], v) <= 1e-12
    assert np.max(np.abs(got["x"] - x) / np.array(prob.dx)[:, None]) <= 1e-11
    sp.destroy(); grid.destroy()


def test_fast_kernel_charge_continuity_large(pgpu):
    """Size-independent property at a larger size: CC1 is charge conserving, so with
    rho deposited by TSC at x_old and x_new, (rho_new - rho_old) + dt * div J = 0 to round-off."""
    from picnic_b200 import decks
    deck = decks.deck_c3(ncell=96, ppc=6, dt=0.1)
    lo, hi = (0, 0), (95, 95)
    E, B = decks.analytic_fields(deck, lo, hi, E0=3.0e7, B0=5.0e8)
    grid = pgpu.Grid(2, deck.ncell, deck.xmin, deck.dx, deck.nghost, (1, 1), volume_scale=1.0)
    grid.set_fields(E, B)
    sdef = deck.species[0]
    p = decks.load_species(deck, sdef, lo, hi, np.random.default_rng(5))
    p["w"][:] = 1.0
    sp = pgpu.Species(grid, sdef.mass, 1.0, sdef.fnorm_const(deck.units), deck.units.cvac_norm, interp_N=1)
    sp.upload(p["x"], p["v"], p["w"])
    rho_old, _, _ = sp.charge_density((1, 1))
    st = sp.advance_iteratively(deck.dt, deposit=True)
    assert st.num_unconverged == 0
    grid.current_zero(); grid.current_add(sp); grid.current_finalize()
    Jx, Jy = grid.current_get(0), grid.current_get(1)
    sp.advance_positions_2nd_half()
    rho_new, _, _ = sp.charge_density((1, 1))
    g = deck.nghost
    n = 96
    # nodes i=0..95 (owned), Jx(c,n): cell i between nodes i and i+1; array offset g
    jx = Jx[g - 1:g + n, g:g + n]         # cells -1..95 at nodes j=0..95
    jy = Jy[g:g + n, g - 1:g + n]
    div = (jx[1:, :] - jx[:-1, :]) / deck.dx[0] + (jy[:, 1:] - jy[:, :-1]) / deck.dx[1]
    drho = (rho_new - rho_old)[g:g + n, g:g + n]
    resid = drho + deck.cnorm_dt * div
    scale = np.max(np.abs(rho_old))
    assert np.max(np.abs(resid)) / scale < 1e-11
    sp.destroy(); grid.destroy()


# ---- the 1D kernel (pgpu_advance_cc1_1d.cu) -------------------------------------------------------------------
def _oracle_1d(prob, rtol=1e-12, itmax=21, charge=-1.0, vs=3.0):
    x, v = prob.x.copy(), prob.v.copy()
    rc, apply_its, unconv, its = orc.advance_particles_iteratively(
        prob.geom, orc.CC1, x, prob.xold, v, prob.vold, prob.E, prob.B, FN, DT * CV, rtol, itmax)
    assert rc == 0
    J0 = prob.new_J()
    assert orc.deposit_current(prob.geom, orc.CC1, x, prob.xold, v, prob.w, DT * CV, J0) == 0
    for c in range(3):
        orc.scale_fab(J0[c], 1, charge / vs)
    return x, v, J0, apply_its, unconv


def _run_1d(pgpu, prob, mode, deposit=True, itmax=21, alias=False):
    pgpu.check(pgpu.load().pgpu_set_deposit_mode(mode))
    grid, sp = make_gpu(pgpu, prob, INTERPS["CC1"], rtol=1e-12, iter_max=itmax, fnorm=FN, cvac_norm=CV,
                        charge=-1.0, volume_scale=3.0)
    if alias:
        sp.update_old_positions(); sp.update_old_velocities()
    pgpu.profile_reset(); pgpu.profile_enable(True)
    st = sp.advance_iteratively(DT, deposit=deposit)
    pgpu.profile_enable(False)
    got = sp.download()
    J = [sp.current_get(c) for c in range(3)] if deposit else None
    nfast = pgpu.profile_query("advance_cc1_1d")[1]
    sp.destroy(); grid.destroy()
    pgpu.check(pgpu.load().pgpu_set_deposit_mode(1))
    return got, J, st, nfast


@pytest.mark.parametrize("order", ["sorted", "shuffled"])
@pytest.mark.parametrize("max_disp", [0.02, 0.5, 1.6])
@pytest.mark.parametrize("n", [20000, 4 * 777 + 3])
def test_fast_kernel_1d_matches_oracle(pgpu, order, max_disp, n):
    prob = Problem(1, (48,), (0.25,), (0.5,), 4, n, seed=25, max_disp=max_disp, E0=0.3, B0=0.8)
    if order == "sorted":
        perm = np.argsort(prob.xold[0], kind="stable")
        for name in ("x", "xold", "v", "vold"):
            setattr(prob, name, np.ascontiguousarray(getattr(prob, name)[:, perm]))
        prob.w = np.ascontiguousarray(prob.w[perm])
    x, v, J0, apply_its, unconv = _oracle_1d(prob)
    got, J, st, nfast = _run_1d(pgpu, prob, 1)
    assert nfast == 1
    assert np.max(np.abs(got["x"] - x)) / prob.dx[0] <= 4e-12
    assert rel_err(got["v"], v) <= 1e-11
    for c in range(3):
        assert rel_err(J[c], J0[c].a) <= 1e-11, c
    assert st.num_parts_its == prob.n and st.num_unconverged == unconv
    assert abs(st.num_apply_its - apply_its) <= max(3, prob.n // 100)
    got0, J00, _, nfast0 = _run_1d(pgpu, prob, 0)
    assert nfast0 == 0
    for c in range(3):
        assert rel_err(J[c], J00[c]) <= 1e-12


def test_fast_kernel_1d_faces_alias_and_single_pass(pgpu):
    """x_old on and next to dual-cell faces (cell centres) and primal faces, inside the ghost region too; then the
    aliased update-old entry (out-of-place write + pointer swap) and advanceParticles (iter_max = 0, no deposit)."""
    ncell, dx, xmin, ng, n = 40, 0.25, 0.5, 4, 9000
    prob = Problem(1, (ncell,), (dx,), (xmin,), ng, n, seed=26, max_disp=0.05, E0=0.3, B0=0.8)
    rng = np.random.default_rng(27)
    k = rng.integers(-2 * (ng - 2), 2 * (ncell + ng - 2), size=n)
    on_face = rng.random(n) < 0.5
    xf = xmin + k * (0.5 * dx)
    eps = rng.choice([-2, -1, 0, 0, 1, 2], size=n)
    nudged = xf.copy()
    for s in (-2, -1, 1, 2):
        m = on_face & (eps == s)
        t = xf[m]
        for _ in range(abs(s)):
            t = np.nextafter(t, np.inf if s > 0 else -np.inf)
        nudged[m] = t
    prob.xold[0] = np.where(on_face, nudged, prob.xold[0])
    prob.x[0] = prob.xold[0] + (rng.random(n) * 2 - 1) * 0.02 * dx
    prob.x[0] = np.where(rng.random(n) < 0.1, prob.xold[0], prob.x[0])
    prob.vold[0] *= 0.05
    prob.v[:] = prob.vold
    x, v, J0, apply_its, unconv = _oracle_1d(prob)
    got, J, st, nfast = _run_1d(pgpu, prob, 1)
    assert nfast == 1
    assert np.max(np.abs(got["x"] - x)) / dx <= 4e-12 and rel_err(got["v"], v) <= 1e-11
    for c in range(3):
        assert rel_err(J[c], J0[c].a) <= 1e-11, c
    # aliased old arrays: the oracle starts from xold = x, vold = v
    prob2 = Problem(1, (ncell,), (dx,), (xmin,), ng, 5001, seed=28, max_disp=0.3, E0=0.3, B0=0.8)
    prob2.xold, prob2.vold = prob2.x.copy(), prob2.v.copy()
    x, v, J0, _, _ = _oracle_1d(prob2)
    got, J, st, nfast = _run_1d(pgpu, prob2, 1, alias=True)
    assert nfast == 1
    assert np.array_equal(got["xold"], prob2.xold) and np.array_equal(got["vold"], prob2.vold)
    assert np.max(np.abs(got["x"] - x)) / dx <= 4e-12 and rel_err(got["v"], v) <= 1e-11
    for c in range(3):
        assert rel_err(J[c], J0[c].a) <= 1e-11, c
    # advanceParticles
    prob3 = Problem(1, (ncell,), (dx,), (xmin,), ng, 7000, seed=29, max_disp=0.1, E0=0.3, B0=0.8)
    got, _, _, nfast = _run_1d(pgpu, prob3, 1, deposit=False, itmax=0)
    assert nfast == 1
    x, v = prob3.x.copy(), prob3.v.copy()
    assert orc.advance_particles(prob3.geom, orc.CC1, x, prob3.xold, v, prob3.vold, prob3.E, prob3.B, FN, DT * CV, 0) == 0
    assert rel_err(got["v"], v) <= 1e-12 and np.max(np.abs(got["x"] - x)) / dx <= 1e-11
