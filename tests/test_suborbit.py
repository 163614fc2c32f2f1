"""The sub-orbit model: PicChargedSpecies::advanceSubOrbitParticlesAndSetJ (PicChargedSpecies.cpp:3324-3669),
transferFastParticles (:894-956), the hand-over of unconverged particles from advanceParticlesIteratively (:1699-1706) and
mergeSubOrbitParticles (:1718-1747).

CPU: the oracle's restatement against what the sub-orbit model IS -- nsub consecutive implicit steps of dt/nsub, each with
its own deposit, the currents averaged -- and against discrete charge continuity.  GPU: the CUDA path against the oracle."""
import numpy as np
import pytest

from common import orc, Problem, make_gpu, rel_err, INTERPS


def _prob(D, seed, n=600):
    if D == 1:
        p = Problem(1, (24,), (0.25,), (0.5,), 4, n, seed=seed, max_disp=0.0, E0=0.3, B0=0.8)
    else:
        p = Problem(2, (12, 10), (0.25, 0.3), (0.5, -1.0), 4, n, seed=seed, max_disp=0.0, E0=0.3, B0=0.8)
    p.x = p.xold.copy()
    p.v = p.vold.copy()
    return p


FN, CDT, RTOL, ITMAX = -0.7, 0.5 * 0.9986, 1e-12, 25


@pytest.mark.parametrize("interp", ["CIC", "CC1"])
@pytest.mark.parametrize("D", [1, 2])
def test_oracle_suborbits_equal_consecutive_short_steps(D, interp):
    """Three sub-orbits == three implicit steps of dt/3 (advanceParticlesIteratively + deposit + 2nd half each), J = mean."""
    prob = _prob(D, 31)
    it = INTERPS[interp]
    nsub = np.full(prob.n, 3, dtype=np.int32)
    x, xold, v, vold = prob.x.copy(), prob.xold.copy(), prob.v.copy(), prob.vold.copy()
    J = prob.new_J()
    rc = orc.advance_suborbit_and_set_J(prob.geom, it, it, x, xold, v, vold, prob.w, nsub, prob.E, prob.B, FN, CDT, RTOL,
                                        ITMAX, J)
    assert rc == 0 and np.all(nsub == 3)
    assert np.array_equal(xold, prob.xold) and np.array_equal(vold, prob.vold)     # restored (:3605-3606)
    xs, vs = prob.xold.copy(), prob.vold.copy()
    Jm = prob.new_J()
    for k in range(3):
        xb, vb = xs.copy(), vs.copy()
        rc, _, unconv, _ = orc.advance_particles_iteratively(prob.geom, it, xb, xs, vb, vs, prob.E, prob.B, FN, CDT / 3, RTOL,
                                                             ITMAX)
        assert rc == 0 and unconv == 0
        Jk = prob.new_J()
        assert orc.deposit_current(prob.geom, it, xb, xs, vb, prob.w, CDT / 3, Jk) == 0
        for c in range(3):
            Jm[c].a += Jk[c].a / 3.0
        xs, vs = 2.0 * xb - xs, 2.0 * vb - vs
    # the sub-orbit loop starts every sub-step from x_bar = x_old (the Picard loop from the stored x_bar): both converge to
    # rtol, so the end states agree to a few rtol
    assert np.abs(x - xs).max() < 20 * RTOL * max(prob.dx) and rel_err(v, vs) < 1e-10
    for c in range(3):
        assert rel_err(J[c].a, Jm[c].a) < 1e-9


@pytest.mark.parametrize("D", [1, 2])
def test_oracle_suborbit_current_satisfies_continuity(D):
    """CC1: rho(x_new) - rho(x_old) + dt div(J_sub) = 0 to round-off, sub-orbit by sub-orbit and hence for their mean."""
    prob = _prob(D, 32, n=200)
    it = INTERPS["CC1"]
    nsub = np.full(prob.n, 2, dtype=np.int32)
    nsub[::3] = 4
    x, xold, v, vold = prob.x.copy(), prob.xold.copy(), prob.v.copy(), prob.vold.copy()
    J = prob.new_J()
    assert orc.advance_suborbit_and_set_J(prob.geom, it, it, x, xold, v, vold, prob.w, nsub, prob.E, prob.B, FN, CDT, RTOL,
                                          ITMAX, J) == 0
    from test_oracle_invariants import _div_J_nodes
    stag = (1,) * D
    rho0 = orc.fab_for(prob.box_lo, prob.box_hi, prob.nghost, stag)
    rho1 = orc.fab_for(prob.box_lo, prob.box_hi, prob.nghost, stag)
    orc.deposit_rho(prob.geom, orc.TSC, prob.xold, prob.w, stag, rho0)      # CC1 current <-> TSC charge
    orc.deposit_rho(prob.geom, orc.TSC, x, prob.w, stag, rho1)
    div, lo = _div_J_nodes(prob, J, CDT)
    if D == 1:
        s0 = lo[0] - rho0.lo[0]
        drho = (rho1.a - rho0.a)[s0:s0 + div.shape[0]]
    else:
        s0, s1 = lo[0] - rho0.lo[0], lo[1] - rho0.lo[1]
        drho = (rho1.a - rho0.a)[s0:s0 + div.shape[0], s1:s1 + div.shape[1]]
    scale = np.abs(rho0.a).max()
    # each sub-orbit's x_bar - x_old equals u_bar cnormDt_sub / 2 to rtol: continuity to a few rtol
    assert np.abs(drho + div).max() / scale < 1e-10
    assert np.abs(drho).max() / scale > 1e-4


def test_oracle_suborbit_adds_suborbits_when_a_substep_does_not_converge():
    """A particle in a field too strong for iter_max passes restarts with one more sub-orbit until it converges (:3521-3541)."""
    prob = _prob(1, 33, n=40)
    it = INTERPS["CC1"]
    nsub = np.full(prob.n, 2, dtype=np.int32)
    x, xold, v, vold = prob.x.copy(), prob.xold.copy(), prob.v.copy(), prob.vold.copy()
    J = prob.new_J()
    rc = orc.advance_suborbit_and_set_J(prob.geom, it, it, x, xold, v, vold, prob.w, nsub, prob.E, prob.B, 40 * FN, CDT, RTOL,
                                        6, J)
    assert rc == 0 and nsub.max() > 2 and nsub.min() >= 2


def test_oracle_fast_particles():
    prob = _prob(2, 34, n=300)
    x = prob.xold.copy()
    x[0, :10] += 0.6 * prob.dx[0]       # x_new = x_old + 1.2 dx: at most 2 crossings  (ghosts - D = 2: allowed)
    x[1, 10:20] += 1.6 * prob.dx[1]     # x_new = x_old + 3.2 dx: 3 or 4 crossings -> fast
    flag = orc.fast_particles(prob.geom, x, prob.xold)
    assert flag[10:20].all() and not flag[:10].any() and not flag[20:].any()


# ---- GPU ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("exact", [1, 0])
@pytest.mark.parametrize("interp", ["CIC", "CC1"])
@pytest.mark.parametrize("D", [1, 2])
def test_gpu_unconverged_particles_go_to_the_suborbit_container_and_match_the_oracle(pgpu, D, interp, exact):
    """iter_max too small for part of the particles: they leave the main container (which deposits without them), take the
    step in sub-orbits, and come back at merge time; everything against the oracle."""
    prob = _prob(D, 35, n=3000)
    it = INTERPS[interp]
    fn, itmax, dt, cv = FN, 10, 0.5, 0.9986
    pgpu.load().pgpu_set_exact_math(exact)
    grid, sp = make_gpu(pgpu, prob, it, rtol=RTOL, iter_max=itmax, fnorm=fn, cvac_norm=cv, charge=-1.0, volume_scale=2.0)
    sp.set_suborbit_model(True)
    st = sp.advance_iteratively(dt, deposit=True)
    nsubc = sp.n_suborbit
    assert 0 < nsubc < prob.n and sp.n == prob.n - nsubc
    # oracle: the same split
    x, v = prob.x.copy(), prob.v.copy()
    rc, _, unconv, its = orc.advance_particles_iteratively(prob.geom, it, x, prob.xold, v, prob.vold, prob.E, prob.B, fn,
                                                           dt * cv, RTOL, itmax)
    stay = its <= itmax
    if exact:
        assert unconv == nsubc
    assert abs(unconv - nsubc) <= max(2, prob.n // 200)
    main = sp.download()
    sub0 = sp.suborbit_download()
    assert np.all(sub0["nsub"] == 2)
    ids_main, ids_sub = main["id"].astype(np.int64), sub0["id"].astype(np.int64)
    assert np.intersect1d(ids_main, ids_sub).size == 0 and ids_main.size + ids_sub.size == prob.n
    both = np.isin(np.arange(prob.n), ids_main) & stay
    o = np.argsort(ids_main)
    sel = ids_main[o]
    keep = both[sel]
    assert np.abs(main["x"][:, o][:, keep] - x[:, sel[keep]]).max() < 4 * RTOL * max(prob.dx) * (1 if exact else 4)
    # J of the main container = deposit of the particles that stayed
    Jm = prob.new_J()
    idm = np.sort(ids_main)
    assert orc.deposit_current(prob.geom, it, np.ascontiguousarray(main["x"][:, o]), np.ascontiguousarray(prob.xold[:, idm]),
                               np.ascontiguousarray(main["v"][:, o]), np.ascontiguousarray(prob.w[idm]), dt * cv, Jm) == 0
    for c in range(3):
        orc.scale_fab(Jm[c], D, -1.0 / 2.0)
        assert rel_err(sp.current_get(c), Jm[c].a) < 1e-11
    # sub-orbit advance + its current
    sp.advance_suborbit_and_set_J(dt)
    sub1 = sp.suborbit_download()
    os_ = np.argsort(ids_sub)
    ids = ids_sub[os_]
    xs, xo = np.ascontiguousarray(prob.xold[:, ids]), np.ascontiguousarray(prob.xold[:, ids])
    vs, vo = np.ascontiguousarray(prob.vold[:, ids]), np.ascontiguousarray(prob.vold[:, ids])
    ws = np.ascontiguousarray(prob.w[ids])
    nsub = np.full(ids.size, 2, dtype=np.int32)
    Js = prob.new_J()
    assert orc.advance_suborbit_and_set_J(prob.geom, it, it, xs, xo, vs, vo, ws, nsub, prob.E, prob.B, fn, dt * cv, RTOL, itmax,
                                          Js) == 0
    assert nsub.min() >= 2
    same = sub1["nsub"][os_] == nsub
    assert same.mean() > (0.999 if exact else 0.98)
    tolx = (1e-14 if exact else 1e-10)
    assert np.abs(sub1["x"][:, os_][:, same] - xs[:, same]).max() < tolx * max(prob.xmax) + 8 * RTOL * max(prob.dx) * (0 if exact else 1)
    assert rel_err(sub1["v"][:, os_][:, same], vs[:, same]) < (1e-14 if exact else 1e-9)
    assert np.array_equal(sub1["xold"][:, os_], prob.xold[:, ids]) and np.array_equal(sub1["vold"][:, os_], prob.vold[:, ids])
    if same.all():
        for c in range(3):
            orc.scale_fab(Js[c], D, -1.0 / 2.0)
            assert rel_err(sp.suborbit_current_get(c), Js[c].a) < (1e-12 if exact else 1e-9)
    # total current and merge
    grid.current_zero(); grid.current_add(sp)
    pgpu.check(pgpu.load().pgpu_current_add_suborbit(grid.h, sp.h))
    tot = [grid.current_get(c) for c in range(3)]
    for c in range(3):
        assert rel_err(tot[c], sp.current_get(c) + sp.suborbit_current_get(c)) < 1e-14
    sp.merge_suborbit()
    assert sp.n == prob.n and sp.n_suborbit == 0
    allp = sp.download()
    assert np.array_equal(np.sort(allp["id"].astype(np.int64)), np.arange(prob.n))
    sp.destroy(); grid.destroy()
    pgpu.load().pgpu_set_exact_math(0)


@pytest.mark.gpu
def test_gpu_suborbit_count_grows_like_the_oracle(pgpu):
    """A field too strong for iter_max passes per sub-step: particles restart with more sub-orbits (:3521-3541); the
    device finds the oracle's counts and end states."""
    prob = _prob(1, 33, n=400)
    it = INTERPS["CIC"]       # no segment limit on the full-step orbits of the first, unconverged attempt
    fn, itmax, dt, cv = 40 * FN, 6, 0.5, 0.9986
    pgpu.load().pgpu_set_exact_math(1)
    grid, sp = make_gpu(pgpu, prob, it, rtol=RTOL, iter_max=itmax, fnorm=fn, cvac_norm=cv, charge=-1.0, volume_scale=2.0)
    sp.set_suborbit_model(True)
    sp.advance_iteratively(dt, deposit=True)
    assert sp.n_suborbit > 50
    sub0 = sp.suborbit_download()
    sp.advance_suborbit_and_set_J(dt)
    sub1 = sp.suborbit_download()
    ids = sub0["id"].astype(np.int64)
    xs, xo = np.ascontiguousarray(prob.xold[:, ids]), np.ascontiguousarray(prob.xold[:, ids])
    vs, vo = np.ascontiguousarray(prob.vold[:, ids]), np.ascontiguousarray(prob.vold[:, ids])
    nsub = np.full(ids.size, 2, dtype=np.int32)
    Js = prob.new_J()
    assert orc.advance_suborbit_and_set_J(prob.geom, it, it, xs, xo, vs, vo, np.ascontiguousarray(prob.w[ids]), nsub, prob.E,
                                          prob.B, fn, dt * cv, RTOL, itmax, Js) == 0
    assert nsub.max() > 2
    same = sub1["nsub"] == nsub
    assert same.mean() > 0.99
    assert np.abs(sub1["x"][:, same] - xs[:, same]).max() < 1e-13 * max(prob.xmax) and rel_err(sub1["v"][:, same], vs[:, same]) < 1e-13
    if same.all():
        for c in range(3):
            orc.scale_fab(Js[c], 1, -1.0 / 2.0)
            assert rel_err(sp.suborbit_current_get(c), Js[c].a) < 1e-12
    sp.destroy(); grid.destroy()
    pgpu.load().pgpu_set_exact_math(0)


@pytest.mark.gpu
def test_gpu_transfer_fast_particles(pgpu):
    prob = _prob(2, 36, n=2000)
    prob.x = prob.xold.copy()
    prob.x[1, 10:20] += 1.6 * prob.dx[1]
    prob.x[0, 40:45] -= 1.7 * prob.dx[0]
    grid, sp = make_gpu(pgpu, prob, INTERPS["CC1"])
    sp.set_suborbit_model(True, fast_particles=True)
    sp.transfer_fast_particles()
    flag = orc.fast_particles(prob.geom, prob.x, prob.xold)
    assert flag.sum() == 15 and sp.n_suborbit == 15 and sp.n == prob.n - 15
    sub = sp.suborbit_download()
    assert np.array_equal(np.sort(sub["id"].astype(np.int64)), np.nonzero(flag)[0]) and np.all(sub["nsub"] == 2)
    sp.destroy(); grid.destroy()
