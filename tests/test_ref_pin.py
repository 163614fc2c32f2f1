"""Pins of the oracle -- and through it of the CUDA path -- on the REFERENCE's own code.

tests/golden/ref_pins.npz holds outputs of the reference's PicSpeciesUtils::applyForces (Boris),
ScatteringUtils::computeDeltaU / rotateVelocity / getScatteringCos and JustinsParticle::linearOut,
produced by tests/golden/make_ref_golden.py from oracle/_ref/libpicnic_ref.so, i.e. from the
reference sources compiled where they lie (oracle/ref_build.sh).  The CPU tests demand BIT equality
of the oracle restatement with those vectors; where /root/reference is present the vectors are also
regenerated live.  The GPU tests drive the same inputs through the C ABI."""
import os

import numpy as np
import pytest

from common import orc, ROOT

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "ref_pins.npz"))
C = lambda a: np.ascontiguousarray(a, dtype=np.float64)


def _oracle_boris(half):
    n = GOLD["in_vold"].shape[1]
    v = np.zeros((3, n))
    orc.lib().orc_boris(n, orc._ptr(v), orc._ptr(C(GOLD["in_vold"])), orc._ptr(C(GOLD["in_Ep"])),
                        orc._ptr(C(GOLD["in_Bp"])), float(GOLD["in_fnorm"]), float(GOLD["in_cnormDt"]), half)
    return v


@pytest.mark.parametrize("half", [0, 1])
def test_oracle_boris_bit_equals_reference(half):
    assert np.array_equal(_oracle_boris(half), GOLD["out_boris_half%d" % half])


def test_oracle_delta_u_bit_equals_reference():
    n = GOLD["in_u"].shape[1]
    got = np.zeros((n, 3))
    lib = orc.lib()
    import ctypes
    lib.orc_scatter_delta_u.argtypes = [ctypes.c_double] * 7 + [ctypes.c_void_p]
    for i in range(n):
        t = np.zeros(3)
        lib.orc_scatter_delta_u(GOLD["in_u"][0, i], GOLD["in_u"][1, i], GOLD["in_u"][2, i], GOLD["in_costh"][i],
                                GOLD["in_sinth"][i], GOLD["in_cosphi"][i], GOLD["in_sinphi"][i], orc._ptr(t))
        got[i] = t
    assert np.array_equal(got, GOLD["out_delta_u"])
    # rotateVelocity(u) == u + computeDeltaU(u) up to round-off: the two reference routines agree,
    # except in the measure-zero branch uperp == 0 with uz < 0, where computeDeltaU (which assumes
    # u is along +z there) does not preserve |u| -- a reference quirk the oracle and the CUDA code keep
    u = GOLD["in_u"]
    quirk = (u[0] == 0.0) & (u[1] == 0.0) & (u[2] < 0.0)
    assert quirk.sum() >= 1 and (~quirk & (u[0] == 0.0) & (u[1] == 0.0)).sum() >= 1
    d = GOLD["out_rotate"] - (u.T + GOLD["out_delta_u"])
    assert np.abs(d[~quirk]).max() <= 4e-16 * np.abs(u).max()


def test_wire_format_layout():
    """JustinsParticle::linearOut in 2D = [w, x0,x1, xold0,xold1, virt0,virt1, v0..2, vold0..2, (double)ID]:
    the 14-double record the migration path of picnic_b200/halo.py carries minus the virtual slots."""
    w = GOLD["out_wire"]
    assert w.shape == (14,)
    assert list(w[:5]) == [7.5, 1.25, -3.5, 1.0, -3.25] and list(w[5:7]) == [0.0, 0.0]
    assert list(w[7:13]) == [0.1, 0.2, 0.3, 0.4, 0.5, 0.6] and w[13] == 123456789.0


def test_golden_regenerates_from_reference():
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("/root/reference not present on this box: committed vectors only")
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "tests", "golden", "make_ref_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    out = mk.run_reference(mk.inputs())
    for k, v in out.items():
        assert np.array_equal(v, GOLD["out_" + k]), k


# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("exact", [1, 0])
@pytest.mark.parametrize("half", [0, 1])
def test_gpu_boris_against_reference_vectors(pgpu, exact, half):
    n = GOLD["in_vold"].shape[1]
    grid = pgpu.Grid(1, (8,), (0.0,), (0.25,), 2, (1,))
    fnorm, cnormDt = float(GOLD["in_fnorm"]), float(GOLD["in_cnormDt"])
    sp = pgpu.Species(grid, 1.0, -1.0, fnorm, 1.0, interp_N=0, interp_J=0, interp_E=0)
    x = np.full((1, n), 1.0)
    sp.upload(x, C(GOLD["in_vold"]), np.ones(n), vold=C(GOLD["in_vold"]))
    sp.set_particle_fields(GOLD["in_Ep"], GOLD["in_Bp"])
    pgpu.check(pgpu.load().pgpu_set_exact_math(exact))
    try:
        sp.advance_velocities(cnormDt, half)     # cvac_norm = 1: full_dt == cnormDt
        got = sp.download()["v"]
    finally:
        pgpu.check(pgpu.load().pgpu_set_exact_math(0))
        sp.destroy(); grid.destroy()
    want = GOLD["out_boris_half%d" % half]
    if exact:
        assert np.array_equal(got, want)                     # reference operation order: bit identical
    else:
        assert np.max(np.abs(got - want)) <= 1e-14 * np.max(np.abs(want))


@pytest.mark.gpu
def test_gpu_delta_u_against_reference_vectors(pgpu):
    got = pgpu.scatter_delta_u(GOLD["in_u"], GOLD["in_costh"], GOLD["in_sinth"], GOLD["in_cosphi"], GOLD["in_sinphi"])
    want = GOLD["out_delta_u"].T
    assert np.max(np.abs(got - want)) <= 4e-15 * np.max(np.abs(GOLD["in_u"]))


# ---- the RELATIVISTIC_PARTICLES build -------------------------------------------------------------
GOLD_REL = np.load(os.path.join(ROOT, "tests", "golden", "ref_pins_rel.npz"))


@pytest.mark.parametrize("hc", [0, 1])
@pytest.mark.parametrize("half", [0, 1])
def test_oracle_relativistic_boris_bit_equals_reference(hc, half):
    """PicSpeciesUtils::applyForces compiled with -DRELATIVISTIC_PARTICLES (Boris gamma and Higuera-Cary)."""
    n = GOLD_REL["in_vold"].shape[1]
    v = np.zeros((3, n))
    orc.set_relativistic(True, bool(hc))
    try:
        orc.lib().orc_boris(n, orc._ptr(v), orc._ptr(C(GOLD_REL["in_vold"])), orc._ptr(C(GOLD_REL["in_Ep"])),
                            orc._ptr(C(GOLD_REL["in_Bp"])), float(GOLD_REL["in_fnorm"]), float(GOLD_REL["in_cnormDt"]),
                            half)
    finally:
        orc.set_relativistic(False)
    assert np.array_equal(v, GOLD_REL["out_boris_hc%d_half%d" % (hc, half)])
    assert not np.array_equal(v, GOLD_REL["out_boris_hc%d_half%d" % (1 - hc, half)])   # the two pushers differ


def test_oracle_implicit_gamma_bit_equals_reference():
    n = GOLD_REL["in_vold"].shape[1]
    got = np.array([orc.implicit_gamma(GOLD_REL["in_vold"][:, i], GOLD_REL["in_ubar"][:, i]) for i in range(n)])
    assert np.array_equal(got, GOLD_REL["out_implicit_gamma"])


def test_relativistic_golden_regenerates_from_reference():
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("no reference checkout on this box")
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "tests", "golden", "make_ref_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    _, out = mk.run_reference_relativistic()
    for k, v in out.items():
        assert np.array_equal(v, GOLD_REL["out_" + k]), k


def test_oracle_rotate_velocity_bit_equals_reference():
    """ScatteringUtils::rotateVelocity (the rotation inside TakizukaAbe::LorentzScatter)."""
    import ctypes
    lib = orc.lib()
    lib.orc_rotate_velocity.argtypes = [ctypes.c_void_p] + [ctypes.c_double] * 4
    n = GOLD["in_u"].shape[1]
    got = np.zeros((n, 3))
    for i in range(n):
        t = np.ascontiguousarray(GOLD["in_u"][:, i].copy())
        lib.orc_rotate_velocity(orc._ptr(t), GOLD["in_costh"][i], GOLD["in_sinth"][i], GOLD["in_cosphi"][i],
                                GOLD["in_sinphi"][i])
        got[i] = t
    assert np.array_equal(got, GOLD["out_rotate"])


def test_oracle_collapse_three_to_two_bit_equals_reference():
    """ScatteringUtils::collapseThreeToTwo (the CONSERVATIVE weight method's 3 -> 2 merge): the oracle against the
    committed outputs of the reference's own code, and live where /root/reference is present; and what it is for:
    total weight, momentum and energy per direction are kept."""
    import ctypes
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "tests", "golden", "make_ref_golden_collapse.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ref_pins_collapse.npz"))
    d = {k[3:]: gold[k] for k in gold.files if k.startswith("in_")}
    f = orc.lib().orc_collapse_three_to_two
    f.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_double]
    o2, o3, w2, w3 = mk.run(f, d)
    for got, name in ((o2, "out_vp2"), (o3, "out_vp3"), (w2, "out_wp2"), (w3, "out_wp3")):
        assert np.array_equal(got, gold[name]), name
    if os.path.isdir("/root/reference"):
        r2, r3, rw2, rw3 = mk.run(mk.ref_lib().ref_collapse_three_to_two, d)
        assert np.array_equal(r2, o2) and np.array_equal(r3, o3) and np.array_equal(rw2, w2)
    wsum0 = d["wp2"] + d["wp3"]
    assert np.allclose(w2 + w3, wsum0, rtol=1e-15)
    p0 = d["wp2p"][:, None] * d["vp2p"] + (d["wp2"] - d["wp2p"])[:, None] * d["vp2"] + d["wp3"][:, None] * d["vp3"]
    p1 = w2[:, None] * o2 + w3[:, None] * o3
    assert np.allclose(p1, p0, rtol=0, atol=1e-15 * np.abs(p0).max())
    e0 = d["wp2p"][:, None] * d["vp2p"] ** 2 + (d["wp2"] - d["wp2p"])[:, None] * d["vp2"] ** 2 + d["wp3"][:, None] * d["vp3"] ** 2
    e1 = w2[:, None] * o2 ** 2 + w3[:, None] * o3 ** 2
    assert np.allclose(e1, e0, rtol=1e-12)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="the reference tree is only in the build container")
@pytest.mark.parametrize("rel", [0, 1])
def test_oracle_mod_energy_pairwise_equals_reference_live(rel):
    """ScatteringUtils::modEnergyPairwise (ScatteringUtils.H:113-205, the energy fix-up of
    scattering.coulomb.enforce_conservations) run from the reference's own code (both builds): velocities to 2 ulp (bit-equal in 9 cases of 10), the two
    long double scalars (remaining deltaE, cumulative relative energy) to 1e-14."""
    import ctypes as C
    so = os.path.join(ROOT, "oracle", "_ref", "libpicnic_ref_rel.so" if rel else "libpicnic_ref.so")
    ref = C.CDLL(so)
    ref.ref_mod_energy_pair.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    lib = orc.lib()
    lib.orc_mod_energy_pairwise.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                            C.c_void_p, C.c_int]
    rng = np.random.default_rng(3)
    moved = 0
    for k in range(300):
        b1, b2 = rng.standard_normal(3) * 0.05, rng.standard_normal(3) * 0.05
        w1, w2 = rng.uniform(0.5, 2), rng.uniform(0.5, 2)
        dE = np.array([rng.standard_normal() * (1e-5 if k % 3 else 1e-9)])
        cum = np.array([rng.uniform(0, 1e-3)])
        a1, a2, dE2, cum2 = b1.copy(), b2.copy(), dE.copy(), cum.copy()
        s1 = b1.copy()
        ref.ref_mod_energy_pair(b1.ctypes.data, b2.ctypes.data, w1, w2, 0.05, cum.ctypes.data, dE.ctypes.data)
        lib.orc_mod_energy_pairwise(a1.ctypes.data, a2.ctypes.data, w1, w2, 0.05, cum2.ctypes.data, dE2.ctypes.data, rel)
        assert np.abs(a1 - b1).max() <= 4e-16 * np.abs(b1).max() and np.abs(a2 - b2).max() <= 4e-16 * np.abs(b2).max()
        # the two long double scalars come back through a double: equal to a few ulp (x87 excess precision differs
        # between the two translation units)
        tol = 1e-12 if rel else 1e-14      # Erel = Ecm - m1 - m2 cancels ten digits in the relativistic form
        assert abs(dE[0] - dE2[0]) <= tol * abs(dE[0]) and abs(cum[0] - cum2[0]) <= tol * cum[0]
        moved += int(not np.array_equal(s1, b1))
    assert moved > 250
