"""The oracle's curvilinear velocity pushes (orc_boris_curvilinear) pinned BIT FOR BIT on the reference's own compiled code:
PicSpeciesUtils::applyForces_CYL_CYL / _SPH_SPH / _CYL_HYB / _SPH_HYB (src/species/pic/PicSpeciesUtils.cpp:103-473), both
builds (-DRELATIVISTIC_PARTICLES too), both byHalfDt, cyclic and anticyclic component order, the predictor-corrector
branch (dtheta == 0) and the stored-angle branch.  Fixtures: tests/golden/ref_pins_curv.npz (make_ref_golden_curv.py);
regenerated live where /root/reference exists.  GPU: the CUDA kernels against the oracle."""
import importlib.util
import os

import numpy as np
import pytest

from common import orc, ROOT

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "ref_pins_curv.npz"))


def _mk():
    spec = importlib.util.spec_from_file_location("mkc", os.path.join(ROOT, "tests", "golden", "make_ref_golden_curv.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    return mk


def _inputs():
    return {k[3:]: GOLD[k] for k in GOLD.files if k.startswith("in_")}


def _args(d):
    c = np.ascontiguousarray
    return c(d["vold"]), c(d["Ep"]), c(d["Bp"]), c(d["r_old"])


def test_oracle_curvilinear_pushes_bit_equal_reference():
    d = _inputs()
    ncase = 0
    for rel, ptype, half, anti in _mk().cases():
        orc.set_relativistic(bool(rel))
        try:
            virt = np.ascontiguousarray(d["virt"].copy())
            v = orc.boris_curvilinear(ptype, *_args(d), virt, float(d["fnorm"]), float(d["cnormDt"]), half, anti)
        finally:
            orc.set_relativistic(False)
        key = "r%d_t%d_h%d_a%d" % (rel, ptype, half, anti)
        assert np.array_equal(v, GOLD["out_v_" + key]), key
        assert np.array_equal(virt, GOLD["out_virt_" + key]), key
        n = virt.shape[1]
        if ptype in (1, 2):      # the predictor-corrector stored an angle where there was none, and left the others
            assert np.all(virt[0, : n // 2] != 0.0) and np.array_equal(virt[:, n // 2:], d["virt"][:, n // 2:])
        else:
            assert np.array_equal(virt, d["virt"])
        ncase += 1
    assert ncase == 18


def test_curvilinear_golden_regenerates_from_reference():
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("/root/reference not present on this box: committed vectors only")
    mk = _mk()
    out = mk.run_reference(mk.inputs())
    for k, v in out.items():
        assert np.array_equal(v, GOLD["out_" + k]), k


def test_cyl_cyl_is_planar_boris_at_large_radius():
    """r -> infinity: the inertia term vanishes and CYL_CYL is applyForces"""
    d = _inputs()
    n = d["vold"].shape[1]
    vold, Ep, Bp, _ = _args(d)
    v = orc.boris_curvilinear(1, vold, Ep, Bp, np.full(n, 1.0e30), np.zeros((2, n)), float(d["fnorm"]), float(d["cnormDt"]), 1)
    ref = orc.boris(np.zeros((3, n)), vold, Ep, Bp, float(d["fnorm"]), float(d["cnormDt"]), 1)
    assert np.max(np.abs(v - ref)) < 1e-16


@pytest.mark.gpu
def test_gpu_curvilinear_pushes_match_reference_vectors(pgpu):
    """pgpu_apply_forces_curvilinear on stored particle fields against the REFERENCE's vectors: every operation is rounded
    as the reference rounds it, so only sin / cos (CUDA vs glibc, 1-2 ulp) can differ."""
    d = _inputs()
    n = d["vold"].shape[1]
    vold, Ep, Bp, r_old = _args(d)
    worst = 0.0
    for rel, ptype, half, anti in _mk().cases():
        grid = pgpu.Grid(1, (64,), (0.0,), (0.25,), 2, (1,))
        sp = pgpu.Species(grid, 1.0, -1.0, float(d["fnorm"]), 1.0, relativistic=bool(rel))
        sp.upload(r_old[None, :], vold, np.ones(n), xold=r_old[None, :], vold=vold)
        sp.set_particle_fields(Ep, Bp)
        sp.set_virtual_positions(d["virt"])
        sp.apply_forces_curvilinear(ptype, float(d["cnormDt"]), half, anti)
        got = sp.download()
        virt = sp.virtual_positions()
        key = "r%d_t%d_h%d_a%d" % (rel, ptype, half, anti)
        ev = np.max(np.abs(got["v"] - GOLD["out_v_" + key])) / np.max(np.abs(GOLD["out_v_" + key]))
        ea = np.max(np.abs(virt - GOLD["out_virt_" + key])) / max(np.max(np.abs(GOLD["out_virt_" + key])), 1e-300)
        assert ev < 1e-15 and ea < 1e-15, (key, ev, ea)
        assert np.array_equal(got["vold"], vold)
        worst = max(worst, ev, ea)
        sp.destroy(); grid.destroy()
    # the planar push is untouched by all this
    assert worst < 1e-15
