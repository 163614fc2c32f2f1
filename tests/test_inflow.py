"""Inflow lists with pic_species.N.suborbit_inflow_J (the implicit shock decks, BASELINE configs[3]):
PicChargedSpecies::advanceInflowParticlesAndSetJ (PicChargedSpecies.cpp:3255-3322) = advanceSubOrbitParticlesAndSetJ with
is_inflow_list (:3376-3669) after advanceInflowPartToBdry (:958-995), then PicChargedSpeciesBC::inflow_Lo / inflow_Hi
(PicChargedSpeciesBC.cpp:961-1001, 1047-1086) in the step's final applyBCs.

CPU: the oracle's restatement against what the model IS -- free streaming to the boundary plane, then the rest of the step
as an ordinary sub-orbit advance from there, current weighted by the time spent inside -- and the turned-around case.
GPU: the CUDA path against the oracle, and the hand-over to the species in pgpu_apply_bcs."""
import numpy as np
import pytest

from common import orc, Problem, make_gpu, rel_err, INTERPS

FN, CDT, RTOL, ITMAX = -0.7, 0.5 * 0.9986, 1e-12, 25


def _prob(D, seed):
    if D == 1:
        return Problem(1, (24,), (0.25,), (0.5,), 4, 10, seed=seed, max_disp=0.0, E0=0.3, B0=0.8)
    return Problem(2, (12, 10), (0.25, 0.3), (0.5, -1.0), 4, 10, seed=seed, max_disp=0.0, E0=0.3, B0=0.8)


def _inflow_particles(prob, n, bdir, side, seed, speed=0.25):
    """As createInflowParticles leaves them: outside the boundary plane by less than what they travel in the step."""
    rng = np.random.default_rng(seed)
    D = prob.D
    v = rng.standard_normal((3, n)) * 0.05
    v[bdir] = (1.0 if side == 0 else -1.0) * (speed * (0.3 + rng.random(n)))
    x = np.zeros((D, n))
    for d in range(D):
        x[d] = prob.xmin[d] + (0.1 + 0.8 * rng.random(n)) * (prob.xmax[d] - prob.xmin[d])
    X0 = prob.xmin[bdir] if side == 0 else prob.xmax[bdir]
    frac = 0.05 + 0.9 * rng.random(n)                       # fraction of the step spent outside
    x[bdir] = X0 - v[bdir] * CDT * frac
    return np.ascontiguousarray(x), np.ascontiguousarray(v), np.ascontiguousarray(rng.random(n) + 0.5), frac


@pytest.mark.parametrize("side", [0, 1])
@pytest.mark.parametrize("interp", ["CIC", "CC1"])
@pytest.mark.parametrize("D", [1, 2])
def test_oracle_inflow_is_free_streaming_plus_suborbits_from_the_boundary(D, interp, side):
    prob = _prob(D, 51)
    it = INTERPS[interp]
    bdir = D - 1
    n = 200
    x0, v0, w, frac = _inflow_particles(prob, n, bdir, side, 52)
    x, xold, v, vold = x0.copy(), x0.copy(), v0.copy(), v0.copy()
    nsub = np.full(n, 1, dtype=np.int32)
    J = prob.new_J()
    rc = orc.advance_inflow_and_set_J(prob.geom, it, it, x, xold, v, vold, w, nsub, prob.E, prob.B, FN, CDT, RTOL, ITMAX, J,
                                      bdir, side)
    assert rc == 0 and np.all(nsub == 1)
    assert np.array_equal(xold, x0) and np.array_equal(vold, v0)          # the original old state is kept (:3605-3606)
    X0 = prob.xmin[bdir] if side == 0 else prob.xmax[bdir]
    Jref = prob.new_J()
    for p in range(n):
        # free streaming to the plane, then ONE ordinary sub-orbit over the rest of the step from there
        dt0 = (X0 - x0[bdir, p]) / v0[bdir, p]
        xb = x0[:, p:p + 1] + v0[:D, p:p + 1] * dt0
        xb[bdir] = X0
        xs, xo = np.ascontiguousarray(xb.copy()), np.ascontiguousarray(xb.copy())
        vs, vo = np.ascontiguousarray(v0[:, p:p + 1].copy()), np.ascontiguousarray(v0[:, p:p + 1].copy())
        one = np.ones(1, dtype=np.int32)
        Jp = prob.new_J()
        assert orc.advance_suborbit_and_set_J(prob.geom, it, it, xs, xo, vs, vo, w[p:p + 1].copy(), one, prob.E, prob.B, FN,
                                              CDT - dt0, RTOL, ITMAX, Jp) == 0
        assert one[0] == 1
        # handed back time-centred against the original old state (:3608-3617)
        assert np.abs(x[:, p] - 0.5 * (xs[:, 0] + x0[:, p])).max() < 1e-14
        assert np.abs(v[:, p] - 0.5 * (vs[:, 0] + v0[:, p])).max() < 1e-15
        # inflow_Lo / Hi will form 2 x - x_old: that is the new-time position, inside the domain
        xn = 2.0 * x[bdir, p] - xold[bdir, p]
        assert (xn >= X0) if side == 0 else (xn < X0)
        for c in range(3):
            Jref[c].a += Jp[c].a * ((CDT - dt0) / CDT)                     # :3655
    for c in range(3):
        assert np.abs(Jref[c].a).max() > 0
        assert rel_err(J[c].a, Jref[c].a) < 1e-13


def test_oracle_inflow_particle_turned_around_before_it_is_inside():
    """A decelerating field strong enough to reverse the particle within the step: it ends the call where it was created,
    with no normal velocity, one sub-orbit and no current (:3486-3509)."""
    prob = _prob(1, 53)
    for f in prob.E:
        f.a[...] = 0.0
    for f in prob.B:
        f.a[...] = 0.0
    prob.E[0].a[...] = 5.0          # alpha E = FN * CDT / 2 * 5 < 0: pushes to -x
    n = 20
    x0, v0, w, frac = _inflow_particles(prob, n, 0, 0, 54, speed=0.05)
    x, xold, v, vold = x0.copy(), x0.copy(), v0.copy(), v0.copy()
    nsub = np.full(n, 2, dtype=np.int32)
    J = prob.new_J()
    assert orc.advance_inflow_and_set_J(prob.geom, orc.CC1, orc.CC1, x, xold, v, vold, w, nsub, prob.E, prob.B, FN, CDT, RTOL,
                                        ITMAX, J, 0, 0) == 0
    assert np.all(nsub == 1)
    assert np.array_equal(x, x0) and np.array_equal(xold, x0) and np.array_equal(vold, v0)
    assert np.all(v[0] == 0.0) and np.array_equal(v[1:], v0[1:])
    for c in range(3):
        assert np.all(J[c].a == 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("exact", [1, 0])
@pytest.mark.parametrize("interp", ["CIC", "CC1"])
@pytest.mark.parametrize("D", [1, 2])
def test_gpu_inflow_advance_and_injection_match_oracle(pgpu, D, interp, exact):
    prob = _prob(D, 55)
    it = INTERPS[interp]
    dt, cv = 0.5, 0.9986
    pgpu.load().pgpu_set_exact_math(exact)
    periodic = [0] * D
    grid, sp = make_gpu(pgpu, prob, it, rtol=RTOL, iter_max=ITMAX, fnorm=FN, cvac_norm=cv, charge=-1.0, volume_scale=2.0,
                        periodic=periodic)
    n0 = sp.n
    lists, tot = [], 0
    for k, (bdir, side) in enumerate([(0, 0), (0, 1)] + ([(1, 0), (1, 1)] if D == 2 else [])):
        x0, v0, w, _ = _inflow_particles(prob, 150 + 10 * k, bdir, side, 60 + k)
        if k == 0:
            v0[bdir, :5] *= 1.0e-3                      # barely moving in: some are turned around by the fields
            x0[bdir, :5] = prob.xmin[bdir] - v0[bdir, :5] * CDT * 0.5
        ids = np.arange(1000 * (k + 1), 1000 * (k + 1) + w.size, dtype=np.uint64)
        sp.inflow_append(x0, v0, w, bdir, side, ids=ids)
        lists.append((bdir, side, x0, v0, w, ids))
        tot += w.size
    assert sp.n_inflow == tot
    sp.advance_inflow_and_set_J(dt)
    got = sp.inflow_download()
    assert np.array_equal(got["boundary"], np.concatenate([np.full(l[4].size, 2 * l[0] + l[1]) for l in lists]))
    Jtot = prob.new_J()
    off, reflected = 0, 0
    expect_join, flux = [], np.zeros((4, 5))
    for (bdir, side, x0, v0, w, ids) in lists:
        n = w.size
        x, xold, v, vold = x0.copy(), x0.copy(), v0.copy(), v0.copy()
        nsub = np.full(n, 1, dtype=np.int32)
        assert orc.advance_inflow_and_set_J(prob.geom, it, it, x, xold, v, vold, w, nsub, prob.E, prob.B, FN, dt * cv, RTOL,
                                            ITMAX, Jtot, bdir, side) == 0
        sl = slice(off, off + n)
        assert np.array_equal(got["nsub"][sl], nsub)
        tol = 1e-14 if exact else 1e-10
        assert np.abs(got["x"][:, sl] - x).max() < tol * max(prob.xmax) and rel_err(got["v"][:, sl], v) < tol
        assert np.array_equal(got["xold"][:, sl], x0) and np.array_equal(got["vold"][:, sl], v0)
        reflected += int(np.sum(v[bdir] == 0.0))
        X0 = prob.xmin[bdir] if side == 0 else prob.xmax[bdir]
        xn = 2.0 * x[bdir] - xold[bdir]
        inside = (xn >= X0) if side == 0 else (xn < X0)
        expect_join.append(ids[inside])
        t = 2 * bdir + side
        flux[t] += [w[inside].sum(), (w * v0[0])[inside].sum(), (w * v0[1])[inside].sum(), (w * v0[2])[inside].sum(),
                    (w * (v0 ** 2).sum(axis=0) / 2.0)[inside].sum()]
        off += n
    assert reflected >= 1
    for c in range(3):
        orc.scale_fab(Jtot[c], D, -1.0 / 2.0)
        assert rel_err(sp.inflow_current_get(c), Jtot[c].a) < (1e-12 if exact else 1e-9)
    # addInflowJ
    grid.current_zero()
    grid.current_add_inflow(sp)
    for c in range(3):
        assert np.array_equal(grid.current_get(c), sp.inflow_current_get(c))
    # the step's final applyBCs: inflow_Lo / inflow_Hi
    bc = (pgpu.BC_INFLOW_OUTFLOW,) * D
    sp.apply_bcs(bc, bc)
    joined = np.concatenate(expect_join)
    assert sp.n_inflow == 0 and sp.n == n0 + joined.size and joined.size == tot - reflected
    allp = sp.download()
    new = allp["id"] >= 1000
    assert np.array_equal(np.sort(allp["id"][new]), np.sort(joined))
    # they carry the new-time state 2 x - x_old and the original old state
    order = {int(i): k for k, i in enumerate(got["id"])}
    idx = np.array([order[int(i)] for i in allp["id"][new]])
    assert np.abs(allp["x"][:, new] - (2.0 * got["x"][:, idx] - got["xold"][:, idx])).max() < 1e-15 * max(np.abs(prob.xmax))
    assert np.abs(allp["v"][:, new] - (2.0 * got["v"][:, idx] - got["vold"][:, idx])).max() < 1e-16
    assert np.array_equal(allp["xold"][:, new], got["xold"][:, idx])
    f = sp.inflow_fluxes()
    assert np.abs(f - flux).max() < 1e-12 * np.abs(flux).max()
    assert np.all(sp.inflow_fluxes() == 0.0)                       # reading resets
    sp.destroy(); grid.destroy()
    pgpu.load().pgpu_set_exact_math(0)


@pytest.mark.gpu
def test_gpu_flow_through_with_inflow_and_outflow_boundaries(pgpu):
    """The life cycle of the shock decks (BASELINE configs[3]) at reduced size, field free: every step the host makes a
    batch of inflow particles outside the left boundary (as createInflowParticles does: where they would be one step
    before they enter), the implicit step advances bulk and inflow lists, the final applyBCs lets the inflow particles
    join and moves the leavers of the right boundary to the outflow list, which the next step removes.
    Checked: exact bookkeeping (injected = inside + left), the steady-state content against the transit times, and --
    the point of the inflow current's cnormDt_sub / cnormDt weight -- a time-averaged J_x that is the same on every
    edge of the domain, the two next to the inflow boundary included."""
    D, ncell, dx, ng = 1, 32, 0.25, 4
    prob = Problem(1, (ncell,), (dx,), (0.0,), ng, 1, seed=3, max_disp=0.0, E0=0.0, B0=0.0)
    for f in prob.E + prob.B:
        f.a[...] = 0.0
    it = INTERPS["CC1"]
    dt, cv = 0.5, 0.9986
    cdt = dt * cv
    grid, sp = make_gpu(pgpu, prob, it, rtol=RTOL, iter_max=ITMAX, fnorm=FN, cvac_norm=cv, charge=-1.0, volume_scale=2.0,
                        periodic=[0])
    sp.upload(np.zeros((1, 0)), np.zeros((3, 0)), np.zeros(0))                  # start empty
    rng = np.random.default_rng(8)
    ud = 0.4 * dx / cdt                                                          # 0.4 cells per step
    bc_lo, bc_hi = (pgpu.BC_INFLOW_OUTFLOW,), (pgpu.BC_OUTFLOW,)
    nsteps, nin, L = 400, 40, ncell * dx
    injected = left = 0.0
    Jsum, nacc, next_id = None, 0, 1
    for step in range(nsteps):
        sp.remove_outflow()
        v = np.zeros((3, nin))
        v[0] = ud * (1.0 + 0.1 * rng.standard_normal(nin)).clip(0.5, 1.5)
        v[1:] = 0.01 * rng.standard_normal((2, nin))
        x = (0.0 - v[0] * cdt * rng.random(nin))[None, :]
        w = np.ones(nin)
        sp.inflow_append(x, v, w, 0, 0, ids=np.arange(next_id, next_id + nin, dtype=np.uint64))
        next_id += nin
        sp.update_old_positions(); sp.update_old_velocities()
        grid.current_zero()
        sp.advance_iteratively(dt, deposit=True, stats=False)
        grid.current_add(sp)
        sp.advance_inflow_and_set_J(dt)
        grid.current_add_inflow(sp)
        grid.current_finalize()
        if step >= nsteps - 150:
            Jx = grid.current_get(0)
            Jsum = Jx.copy() if Jsum is None else Jsum + Jx
            nacc += 1
        sp.advance_velocities_2nd_half(); sp.advance_positions_2nd_half()
        sp.apply_bcs(bc_lo, bc_hi)
        assert sp.n_inflow == 0
        injected += sp.inflow_fluxes()[0, 0]
        left += sp.outflow_fluxes()[1, 0]
    assert injected == nsteps * nin                                              # nothing is turned around without fields
    assert injected == sp.n + left                                               # exact bookkeeping
    # steady state: a particle stays L / (v cdt) steps; <1/v> over the clipped normal factor
    f = (1.0 + 0.1 * np.random.default_rng(1).standard_normal(400000)).clip(0.5, 1.5)
    transit = L / (ud * cdt) * np.mean(1.0 / f)
    assert abs(sp.n / (nin * transit) - 1.0) < 0.03
    all_p = sp.download()
    assert np.all((all_p["x"][0] >= 0.0) & (all_p["x"][0] < L))
    # J_x on the cell-centred edges 0 .. ncell-1 of the domain (array index ng + k)
    J = Jsum[ng:ng + ncell] / nacc
    assert np.all(J < 0.0)                                                       # electrons moving to +x
    interior = J[4:-4].mean()
    # every edge but the two boundary ones carries the same current ...
    assert np.abs(J[1:-1] / interior - 1.0).max() < 0.01, J / interior
    # ... and at the inflow face the CC1 shape puts 1/8 of it on the ghost edge outside (a particle in the first half
    # cell shares its current between edges -1 and 0): together they carry all of it.  A wrong time weight of the inflow
    # current (the particles spend only part of the step inside) would show here at the 50 % level.
    Jm1 = Jsum[ng - 1] / nacc
    assert abs((Jm1 + J[0]) / interior - 1.0) < 0.005 and abs(Jm1 / interior - 0.125) < 0.005
    assert Jsum[ng - 2] == 0.0
    # and its value: charge * (weight per step) / (cnormDt * volume_scale) through every face
    assert abs(interior / (-1.0 * nin / cdt / 2.0) - 1.0) < 0.02
    sp.destroy(); grid.destroy()
