"""The fused PIC_EM_EXPLICIT particle step (pgpu_explicit_step, k_explicit_step): one pass over the particles against
(a) the same step composed of the separate reference-named calls and (b) the oracle's restatement of that sequence
(PICTimeIntegrator_EM_Explicit.cpp:92-170: interpolateFieldsToParticles, advanceVelocities(dt, false),
advancePositionsExplicit(dt/2), applyBCs, setCurrentDensity(dt, true), advancePositions_2ndHalf, applyBCs)."""
import numpy as np
import pytest

from common import orc, Problem, make_gpu, rel_err, INTERPS


def _prob(D, seed):
    # explicit start of step: x_old == x and u_old == u (updateOldParticlePositions / Velocities)
    if D == 1:
        p = Problem(1, (24,), (0.25,), (0.5,), 4, 5000, seed=seed, max_disp=0.0)
    else:
        p = Problem(2, (12, 10), (0.25, 0.3), (0.5, -1.0), 4, 5000, seed=seed, max_disp=0.0)
    p.x = p.xold.copy()
    p.v = p.vold.copy()
    return p


def _separate(sp, dt, bc, second_half):
    sp.interpolate_fields()
    sp.advance_velocities(dt, False)
    sp.advance_positions_explicit(dt, half=True)
    sp.apply_bcs(bc, bc)
    sp.set_current_density(dt, from_explicit=True)
    if second_half:
        sp.advance_positions_2nd_half()
        sp.apply_bcs(bc, bc)


@pytest.mark.gpu
@pytest.mark.parametrize("second_half", [False, True])
@pytest.mark.parametrize("interp", ["CIC", "TSC", "CC0", "CC1"])
@pytest.mark.parametrize("D", [1, 2])
def test_fused_explicit_step_equals_separate_calls(pgpu, D, interp, second_half):
    prob = _prob(D, 21)
    fn, dt, cv = 0.08, 0.9, 1.0
    bc = (1,) * D
    pgpu.load().pgpu_set_exact_math(1)
    out, Js = [], []
    for fused in (False, True):
        grid, sp = make_gpu(pgpu, prob, INTERPS[interp], fnorm=fn, cvac_norm=cv, charge=-1.0, volume_scale=2.0)
        if fused:
            sp.explicit_step(dt, bc, bc, second_half)
        else:
            _separate(sp, dt, bc, second_half)
        out.append(sp.download())
        Js.append([sp.current_get(c) for c in range(3)])
        sp.destroy(); grid.destroy()
    pgpu.load().pgpu_set_exact_math(0)
    # the same operations per particle in the same order: identical particles; J differs by the order of the atomics
    for k in ("x", "xold", "v"):
        assert np.array_equal(out[0][k], out[1][k]), k
    for c in range(3):
        assert rel_err(Js[1][c], Js[0][c]) < 2e-14
    # particles did move and wrap
    assert np.abs(out[1]["x"] - prob.x).max() > 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("interp", ["CIC", "TSC"])
@pytest.mark.parametrize("D", [1, 2])
def test_fused_explicit_step_matches_oracle(pgpu, D, interp):
    prob = _prob(D, 22)
    fn, dt, cv = 0.08, 0.9, 1.0
    it = INTERPS[interp]
    grid, sp = make_gpu(pgpu, prob, it, fnorm=fn, cvac_norm=cv, charge=-1.0, volume_scale=2.0)
    sp.explicit_step(dt, (1,) * D, (1,) * D, True)
    got = sp.download()
    J = [sp.current_get(c) for c in range(3)]
    n = prob.n
    rc, Ep, Bp = orc.gather(prob.geom, it, prob.x, prob.xold, prob.E, prob.B)
    assert rc == 0
    v = np.ascontiguousarray(orc.boris(np.zeros((3, n)), prob.vold, Ep, Bp, fn, dt * cv, 0))
    x = prob.x.copy()
    xold = prob.xold.copy()
    orc.lib().orc_advance_positions_explicit(D, n, orc._ptr(x), orc._ptr(xold), orc._ptr(v), cv * dt * 0.5)
    for d in range(D):
        orc.lib().orc_bc_periodic(n, x[d].ctypes.data, xold[d].ctypes.data, prob.xmin[d], prob.xmax[d])
    J0 = prob.new_J()
    assert orc.deposit_current(prob.geom, it, x, xold, v, prob.w, dt * cv, J0) == 0
    orc.lib().orc_advance_positions_2nd_half(D, n, orc._ptr(x), orc._ptr(xold))
    for d in range(D):
        orc.lib().orc_bc_periodic(n, x[d].ctypes.data, xold[d].ctypes.data, prob.xmin[d], prob.xmax[d])
    assert rel_err(got["v"], v) < 1e-12
    assert np.abs(got["x"] - x).max() < 1e-12 * max(prob.xmax)
    for c in range(3):
        orc.scale_fab(J0[c], D, -1.0 / 2.0)
        assert rel_err(J[c], J0[c].a) < 1e-12
    sp.destroy(); grid.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("second_half", [False, True])
@pytest.mark.parametrize("order", ["sorted", "shuffled"])
def test_explicit_step_cc1_tile_kernel_matches_oracle(pgpu, order, second_half):
    """2D CC1/CC1 in fast arithmetic: the step goes through the tile kernel of the implicit advance (single pass,
    u_new = 2 ubar - u_old) and the one-pass visitor kernel for the particles it defers.  Against the oracle's restatement
    of the separate reference calls.  The tile kernel deposits before the periodic wrap (the generic kernel after it), so
    the currents are compared after the ghost fold, which is what the field solve sees."""
    D, it = 2, INTERPS["CC1"]
    prob = Problem(2, (24, 20), (0.25, 0.3), (0.5, -1.0), 4, 20000, seed=41, max_disp=0.0)
    prob.x = prob.xold.copy()
    prob.v = prob.vold.copy()
    if order == "sorted":
        cells = orc.bin_cells(prob.geom, prob.xold)
        perm = np.argsort(cells[0] + cells[1] * prob.ncell[0], kind="stable")
        for name in ("x", "xold", "v", "vold"):
            setattr(prob, name, np.ascontiguousarray(getattr(prob, name)[:, perm]))
        prob.w = np.ascontiguousarray(prob.w[perm])
    fn, dt, cv = 0.08, 0.9, 1.0
    grid, sp = make_gpu(pgpu, prob, it, fnorm=fn, cvac_norm=cv, charge=-1.0, volume_scale=2.0)
    pgpu.profile_reset(); pgpu.profile_enable(True)
    sp.explicit_step(dt, (1, 1), (1, 1), second_half)
    pgpu.profile_enable(False)
    assert pgpu.profile_query("explicit_step_cc1")[1] == 1          # the tile kernel ran
    assert pgpu.profile_query("explicit_step_deferred")[1] == 1     # and the visitor kernel took its list
    assert pgpu.profile_query("explicit_step_fused")[1] == 0
    got = sp.download()
    J = [sp.current_get(c) for c in range(3)]
    n = prob.n
    rc, Ep, Bp = orc.gather(prob.geom, it, prob.x, prob.xold, prob.E, prob.B)
    assert rc == 0
    v = np.ascontiguousarray(orc.boris(np.zeros((3, n)), prob.vold, Ep, Bp, fn, dt * cv, 0))
    x = prob.x.copy()
    xold = prob.xold.copy()
    orc.lib().orc_advance_positions_explicit(D, n, orc._ptr(x), orc._ptr(xold), orc._ptr(v), cv * dt * 0.5)
    for d in range(D):
        orc.lib().orc_bc_periodic(n, x[d].ctypes.data, xold[d].ctypes.data, prob.xmin[d], prob.xmax[d])
    J0 = prob.new_J()
    assert orc.deposit_current(prob.geom, it, x, xold, v, prob.w, dt * cv, J0) == 0
    if second_half:
        orc.lib().orc_advance_positions_2nd_half(D, n, orc._ptr(x), orc._ptr(xold))
        for d in range(D):
            orc.lib().orc_bc_periodic(n, x[d].ctypes.data, xold[d].ctypes.data, prob.xmin[d], prob.xmax[d])
    assert rel_err(got["v"], v) < 1e-12
    assert np.abs(got["x"] - x).max() < 1e-12 * max(prob.xmax)
    assert np.abs(got["xold"] - xold).max() < 1e-12 * max(prob.xmax)
    assert np.abs(got["x"] - prob.x).max() > 1e-3                   # particles did move
    lo, hi = (0, 0), tuple(nc - 1 for nc in prob.ncell)
    for c, stag in enumerate(orc.E_STAG[2]):
        orc.scale_fab(J0[c], D, -1.0 / 2.0)
        Jg = orc.Fab(J0[c].lo, J0[c].hi, np.asfortranarray(J[c]))
        orc.fold_periodic(J0[c], D, stag, lo, hi, (1, 1))
        orc.fold_periodic(Jg, D, stag, lo, hi, (1, 1))
        assert np.abs(J0[c].a).max() > 0
        assert rel_err(Jg.a, J0[c].a) < 1e-12
    sp.destroy(); grid.destroy()
