"""Outflow / inflow boundaries on the device (second slice): PicChargedSpeciesBC::apply with "outflow" /
"inflow_outflow" boundaries (PicChargedSpeciesBC.cpp:187-224: outflow_Lo/Hi, :872-918), depositInflowOutflowJ in the explicit
solver's setCurrentDensity (:667-736; PicChargedSpecies.cpp:3232-3235), the flux diagnostics, removeOutflowParticles
(:508-545) and the injection of host-made inflow particles (injectInflowParticles, :563-665)."""
import numpy as np
import pytest

from common import orc, Problem, make_gpu, rel_err, INTERPS

OUTFLOW, INFLOW_OUTFLOW, PERIODIC, NONE = 3, 4, 1, 0


def _prob(D, seed):
    if D == 1:
        p = Problem(1, (24,), (0.25,), (0.5,), 4, 4000, seed=seed, max_disp=0.0)
    else:
        p = Problem(2, (12, 10), (0.25, 0.3), (0.5, -1.0), 4, 4000, seed=seed, max_disp=0.0)
    # one explicit half push has happened: x = x_old + u dt/2, part of the particles are now beyond the domain
    p.v = p.vold.copy()
    p.v[0] *= 7.0             # fast along the direction of the outflow boundaries only: the lists keep their un-wrapped
                              # transverse position and must stay inside the ghost layers to deposit
    p.x = np.ascontiguousarray(p.xold + p.v[:D] * 0.35)
    return p


@pytest.mark.gpu
@pytest.mark.parametrize("interp", ["CIC", "TSC"])
@pytest.mark.parametrize("D", [1, 2])
def test_outflow_lists_fluxes_current_and_removal(pgpu, D, interp):
    prob = _prob(D, 41)
    it = INTERPS[interp]
    grid, sp = make_gpu(pgpu, prob, it, charge=-1.0, volume_scale=2.0, periodic=[0] * D)
    bc_lo = (OUTFLOW,) + ((PERIODIC,) if D == 2 else ())
    bc_hi = (INFLOW_OUTFLOW,) + ((PERIODIC,) if D == 2 else ())
    sp.apply_bcs(bc_lo, bc_hi)
    lo_out = prob.x[0] < prob.xmin[0]
    hi_out = prob.x[0] >= prob.xmax[0]
    nout = int(lo_out.sum() + hi_out.sum())
    assert 50 < nout < prob.n // 2 and sp.n_outflow == nout and sp.n == prob.n - nout
    out = sp.outflow_download()
    ids = out["id"].astype(np.int64)
    assert np.array_equal(np.sort(ids), np.nonzero(lo_out | hi_out)[0])
    assert np.array_equal(out["boundary"], np.where(lo_out[ids], 0, 1))
    for k in ("x", "v", "xold", "vold"):
        assert np.array_equal(out[k], getattr(prob, k)[:, ids])
    main = sp.download()
    idm = main["id"].astype(np.int64)
    assert np.array_equal(np.sort(idm), np.nonzero(~(lo_out | hi_out))[0])
    if D == 2:      # the periodic direction wrapped the stayers
        assert np.all((main["x"][1] >= prob.xmin[1]) & (main["x"][1] < prob.xmax[1]))
    # flux diagnostics per boundary: sums of w, w u_old, w |u_old|^2 / 2
    fl = sp.outflow_fluxes()
    for b, m in ((0, lo_out), (1, hi_out)):
        w, u = prob.w[m], prob.vold[:, m]
        ref = [w.sum(), (w * u[0]).sum(), (w * u[1]).sum(), (w * u[2]).sum(), (w * (u ** 2).sum(0)).sum() / 2.0]
        assert np.allclose(fl[b], ref, rtol=1e-12, atol=1e-12 * abs(ref[0]))
    assert not fl[2:].any()
    # explicit solver's setCurrentDensity: main container + outflow lists (depositInflowOutflowJ)
    sp.set_current_density(0.7, from_explicit=True)
    J0 = prob.new_J()
    order = np.concatenate([idm, ids])
    xs = np.concatenate([main["x"], out["x"]], axis=1)
    xo = np.concatenate([main["xold"], out["xold"]], axis=1)
    assert orc.deposit_current(prob.geom, it, np.ascontiguousarray(xs), np.ascontiguousarray(xo),
                               np.ascontiguousarray(prob.v[:, order]), np.ascontiguousarray(prob.w[order]), 0.7, J0) == 0
    for c in range(3):
        orc.scale_fab(J0[c], D, -1.0 / 2.0)
        assert rel_err(sp.current_get(c), J0[c].a) < 1e-12
    # implicit path (from_explicit_solver = false) leaves the lists out
    sp.set_current_density(0.7, from_explicit=False)
    J1 = prob.new_J()
    assert orc.deposit_current(prob.geom, it, np.ascontiguousarray(main["x"]), np.ascontiguousarray(main["xold"]),
                               np.ascontiguousarray(prob.v[:, idm]), np.ascontiguousarray(prob.w[idm]), 0.7, J1) == 0
    for c in range(3):
        orc.scale_fab(J1[c], D, -1.0 / 2.0)
        assert rel_err(sp.current_get(c), J1[c].a) < 1e-12
    sp.remove_outflow()
    assert sp.n_outflow == 0
    # inflow: host-made particles enter the main container
    rng = np.random.default_rng(5)
    m = 37
    xi = np.array(prob.xmin)[:, None] + rng.random((D, m)) * 0.1
    vi = rng.standard_normal((3, m)) * 0.01
    wi = rng.random(m) + 0.5
    sp.append(xi, vi, wi, ids=np.arange(m, dtype=np.uint64) + 10 ** 6)
    assert sp.n == prob.n - nout + m
    after = sp.download()
    tail = after["id"] >= 10 ** 6
    assert tail.sum() == m and np.array_equal(after["x"][:, tail], xi) and np.array_equal(after["vold"][:, tail], vi)
    sp.destroy(); grid.destroy()
