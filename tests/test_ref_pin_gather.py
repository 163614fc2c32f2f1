"""Pins of the oracle's gather restatement -- and, through adjointness, of its deposit weights -- on the REFERENCE's
own C++ gathers.

tests/golden/ref_pins_gather.npz holds E_p / B_p produced by the reference's MeshInterp::interpolateEMfieldsToPart_testing
(-> interpolateEMfieldsToPart_CIC / _TSC, interpolateBfieldsToPart_CIC, interpolateEToPart_CC0 in 1D and 2D,
interpolateEToPart_CC1 in 1D; src/particle_tools/MeshInterpI.H:1013-1850), compiled from where it lies under
/root/reference (oracle/ref_build.sh, oracle/ref_meshinterp.cpp) and run by tests/golden/make_ref_golden_gather.py.

Those routines are the author's C++ statement of the gathers the production build runs from the Fortran
(MeshInterpF.ChF, MeshInterpChargeConservingF.ChF), which the oracle follows operation by operation: the same stencil
points, segment walk and weights in a different order of operations (e.g. (le + (i + 1/2) dx - x)/dx against
((i dx + dx/2) - x + le)/dx).  So the bar is not bit equality but agreement to a few units of round-off of the stencil sum --
any wrong index, weight, stagger or segment split is an O(1) error.  The deposit of each shape uses the identical weights
(checked to round-off by the adjointness tests in test_oracle_invariants.py), so this pins both.

What stays restatement-only: the 2D CC1 weights (the reference's 2D C++ CC1 routine is, in its own words, "just a copy of
_CC0 in 2D" and is never dispatched); they rest on the pinned 1D CC1 weights, the pinned 2D CC0 segment walk (the same walk on
the half-shifted grid) and the continuity / adjointness invariants."""
import os

import numpy as np
import pytest

from common import orc, ROOT

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "ref_pins_gather.npz"))
import importlib.util
_spec = importlib.util.spec_from_file_location("make_ref_golden_gather",
                                               os.path.join(ROOT, "tests", "golden", "make_ref_golden_gather.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)

ORC_INTERP = {"CIC": orc.CIC, "TSC": orc.TSC, "CC0": orc.CC0, "CC1": orc.CC1}
EPS = np.finfo(np.float64).eps


def _oracle(D, interp, tag):
    le = gen.XMIN[:D]
    dx = gen.DX[:D]
    re = [l + gen.NCELL * h for l, h in zip(le, dx)]
    geom = orc.make_geom(D, le, re, dx, gen.GHOSTS)
    lo_box, hi_box = [0] * D, [gen.NCELL - 1] * D
    fabs = []
    for c, stag in enumerate(gen.E_STAG[D] + gen.B_STAG[D]):
        lo = [l - gen.GHOSTS for l in lo_box]
        hi = [h + gen.GHOSTS + s for h, s in zip(hi_box, stag)]
        fabs.append(orc.Fab(lo, hi, GOLD["F%d_%s" % (c, tag)]))
    x = np.ascontiguousarray(GOLD["x_" + tag])
    xold = np.ascontiguousarray(GOLD["xold_" + tag])
    rc, Ep, Bp = orc.gather(geom, ORC_INTERP[interp], x, xold, fabs[:3], fabs[3:])
    assert rc == 0
    return Ep, Bp


@pytest.mark.parametrize("D,interp", gen.CASES)
def test_oracle_gather_matches_reference_cpp(D, interp):
    tag = "%dd_%s" % (D, interp)
    Ep, Bp = _oracle(D, interp, tag)
    # The reference's C++ CC0 walk divides dXp_sub by dXp without the guard its Fortran twin has (seg_factor = 1 where
    # dXp == 0, MeshInterpChargeConservingF.ChF:281-289): an orbit along one axis that crosses a cell face gives 0/0 = NaN
    # in the other in-plane component.  The oracle follows the Fortran; those entries are excluded, and must be exactly those.
    x, xold = GOLD["x_" + tag], GOLD["xold_" + tag]
    axis_aligned = np.any(x == xold, axis=0) & np.any(x != xold, axis=0)
    for got, name, base in ((Ep, "Ep_", 0), (Bp, "Bp_", 3)):
        ref = GOLD[name + tag]
        for c in range(3):
            scale = np.abs(GOLD["F%d_%s" % (base + c, tag)]).max()
            nan = np.isnan(ref[c])
            assert not np.any(nan & ~axis_aligned) and np.all(np.isfinite(got[c])), (tag, name, c)
            assert nan.sum() <= 8 and (not nan.any() or (interp == "CC0" and D == 2 and c < 2))
            err = np.abs(got[c] - ref[c])[~nan].max()
            # a 2D TSC stencil sums nine products; the CC gathers up to three segments of six: <= 16 ulp of the scale
            assert err <= 16 * EPS * scale, (tag, name, c, err / (EPS * scale))
    # the pins are not vacuous: fields are O(1), particles see different values
    assert np.nanmax(np.abs(GOLD["Ep_" + tag])) > 0.5 and np.nanmax(GOLD["Ep_" + tag][0]) - np.nanmin(GOLD["Ep_" + tag][0]) > 0.5


@pytest.mark.parametrize("D,interp", [(1, "CC0"), (1, "CC1"), (2, "CC0")])
def test_pins_cover_multi_segment_orbits(D, interp):
    """The charge-conserving pins exercise the segment walk: orbits with 2 and 3 segments are present."""
    tag = "%dd_%s" % (D, interp)
    x, xold = GOLD["x_" + tag], GOLD["xold_" + tag]
    le = np.array(gen.XMIN[:D])[:, None]
    dx = np.array(gen.DX[:D])[:, None]
    shift = 0.5 if interp == "CC1" else 0.0
    xnew = 2.0 * x - xold
    cross = np.abs(np.floor((xnew - le) / dx - shift) - np.floor((xold - le) / dx - shift)).sum(axis=0)
    assert (cross == 0).sum() > 20 and (cross == 1).sum() > 20 and (cross >= 2).sum() > 5


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="the reference tree is only in the build container")
@pytest.mark.parametrize("D,interp", gen.CASES)
def test_golden_regenerates_from_reference(D, interp):
    """Where /root/reference is present the committed vectors are regenerated live, bit for bit."""
    tag = "%dd_%s" % (D, interp)
    fields, x, xold = gen.inputs(D, 1000 * D + gen.REF_INTERP[interp])
    assert np.array_equal(x, GOLD["x_" + tag]) and np.array_equal(xold, GOLD["xold_" + tag])
    Ep, Bp = gen.run_reference(D, interp, fields, x, xold)
    assert np.array_equal(Ep, GOLD["Ep_" + tag], equal_nan=True) and np.array_equal(Bp, GOLD["Bp_" + tag])


@pytest.mark.gpu
@pytest.mark.parametrize("exact", [1, 0])
@pytest.mark.parametrize("D,interp", gen.CASES)
def test_gpu_gather_matches_reference_cpp(pgpu, D, interp, exact):
    """The CUDA gather (pgpu_interpolate_fields_to_particles, through the C ABI) on the reference's own vectors: within
    1e-12 of the field scale of what the reference's C++ routine returned (north-star bar for the fp64 path)."""
    tag = "%dd_%s" % (D, interp)
    le, dx = gen.XMIN[:D], gen.DX[:D]
    pgpu.load().pgpu_set_exact_math(exact)
    grid = pgpu.Grid(D, (gen.NCELL,) * D, le, dx, gen.GHOSTS, (0,) * D)
    lo_box, hi_box = [0] * D, [gen.NCELL - 1] * D
    comps = []
    for c, stag in enumerate(gen.E_STAG[D] + gen.B_STAG[D]):
        lo = [l - gen.GHOSTS for l in lo_box]
        hi = [h + gen.GHOSTS + s for h, s in zip(hi_box, stag)]
        comps.append((lo, hi, np.asfortranarray(GOLD["F%d_%s" % (c, tag)])))
    grid.set_fields(comps[:3], comps[3:])
    it = ORC_INTERP[interp]
    sp = pgpu.Species(grid, 1.0, -1.0, 1.0, 1.0, interp_N=orc.TSC, interp_J=it, interp_E=it)
    x, xold = np.ascontiguousarray(GOLD["x_" + tag]), np.ascontiguousarray(GOLD["xold_" + tag])
    n = x.shape[1]
    sp.upload(x, np.zeros((3, n)), np.ones(n), xold=xold)
    sp.interpolate_fields()
    Ep, Bp = sp.particle_fields()
    for got, name, base in ((Ep, "Ep_", 0), (Bp, "Bp_", 3)):
        ref = GOLD[name + tag]
        for c in range(3):
            scale = np.abs(GOLD["F%d_%s" % (base + c, tag)]).max()
            ok = ~np.isnan(ref[c])
            assert np.abs(got[c] - ref[c])[ok].max() <= 1e-12 * scale, (tag, name, c)
    sp.destroy(); grid.destroy()
    pgpu.load().pgpu_set_exact_math(0)
