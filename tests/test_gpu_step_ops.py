"""Step-level device ops: the lazy gather of x_old/v_old after a cell sort, and the fused
2nd-half + periodic-BC kernel, each against the separate reference-ordered calls."""
import numpy as np
import pytest

from common import INTERPS, Problem, make_gpu, orc

pytestmark = pytest.mark.gpu


def _prob(seed=31, n=5000):
    return Problem(2, (12, 10), (0.25, 0.3), (0.5, -1.0), 4, n, seed=seed, max_disp=0.4)


def _inside(prob):
    for d in range(2):
        prob.x[d] = np.clip(prob.x[d], prob.xmin[d], prob.xmax[d] - 1e-9 * prob.dx[d])


@pytest.mark.parametrize("reader", ["download", "update_old", "second_half", "advance", "sort_again", "bcs"])
def test_sort_keeps_old_arrays_consistent(pgpu, reader):
    """After bin_particles the old arrays are gathered lazily; whatever touches them next must
    see them in the sorted order (same particle <-> same id)."""
    prob = _prob()
    _inside(prob)
    grid, sp = make_gpu(pgpu, prob, INTERPS["CC1"], fnorm=-0.7, cvac_norm=0.9986)
    sp.bin_particles()
    if reader == "update_old":
        sp.update_old_positions(); sp.update_old_velocities()
        got = sp.download(); o = got["id"].astype(np.int64)
        assert np.array_equal(got["xold"], prob.x[:, o]) and np.array_equal(got["vold"], prob.v[:, o])
    elif reader == "second_half":
        sp.advance_positions_2nd_half(); sp.advance_velocities_2nd_half()
        got = sp.download(); o = got["id"].astype(np.int64)
        assert np.array_equal(got["x"], 2.0 * prob.x[:, o] - prob.xold[:, o])
        assert np.array_equal(got["v"], 2.0 * prob.v[:, o] - prob.vold[:, o])
    elif reader == "advance":
        st = sp.advance_iteratively(0.5, deposit=True)
        got = sp.download(); o = got["id"].astype(np.int64)
        x, v = prob.x.copy(), prob.v.copy()
        rc, _, _, _ = orc.advance_particles_iteratively(prob.geom, orc.CC1, x, prob.xold, v, prob.vold, prob.E, prob.B,
                                                        -0.7, 0.5 * 0.9986, 1e-12, 21)
        assert rc == 0
        assert np.max(np.abs(got["x"] - x[:, o]) / np.array(prob.dx)[:, None]) <= 4e-12
        assert np.max(np.abs(got["v"] - v[:, o])) / np.max(np.abs(v)) <= 1e-11
    elif reader == "sort_again":
        sp.update_old_positions()            # drops the pending position gather only
        sp.bin_particles()                   # must first apply the pending velocity gather
        got = sp.download(); o = got["id"].astype(np.int64)
        assert np.array_equal(got["vold"], prob.vold[:, o]) and np.array_equal(got["xold"], prob.x[:, o])
    elif reader == "bcs":
        sp.apply_bcs((1, 1), (1, 1))
    got = sp.download(); o = got["id"].astype(np.int64)
    assert np.array_equal(np.sort(o), np.arange(prob.n))
    if reader in ("download", "bcs"):
        for name in ("x", "xold", "v", "vold"):
            assert np.array_equal(got[name], getattr(prob, name)[:, o]), name
        assert np.array_equal(got["w"], prob.w[o])
    sp.destroy(); grid.destroy()


@pytest.mark.parametrize("D", [1, 2])
def test_finish_implicit_step_equals_separate_calls(pgpu, D):
    if D == 1:
        prob = Problem(1, (24,), (0.25,), (0.5,), 4, 4000, seed=33, max_disp=1.5)
    else:
        prob = Problem(2, (12, 10), (0.25, 0.3), (0.5, -1.0), 4, 4000, seed=33, max_disp=1.5)
    # xbar near/over the domain edges so that x = 2 xbar - xold leaves the domain on both sides
    per = (1,) * D
    res = []
    for fused in (0, 1):
        grid, sp = make_gpu(pgpu, prob, INTERPS["CIC"])
        if fused:
            sp.finish_implicit_step(per, per)
        else:
            sp.advance_velocities_2nd_half(); sp.advance_positions_2nd_half(); sp.apply_bcs(per, per)
        res.append(sp.download())
        sp.destroy(); grid.destroy()
    for name in ("x", "xold", "v", "vold"):
        assert np.array_equal(res[0][name], res[1][name]), name
    L = np.array(prob.xmax) - np.array(prob.xmin)
    assert np.all(res[1]["x"] >= np.array(prob.xmin)[:, None]) and np.all(res[1]["x"] < np.array(prob.xmax)[:, None])
    moved = np.abs(res[1]["x"] - (2.0 * prob.x - prob.xold)) > 0.5 * L[:, None]
    assert moved.any()                                   # the wrap was exercised
    # symmetry walls take the unfused route and still agree with the oracle's order of operations
    grid, sp = make_gpu(pgpu, prob, INTERPS["CIC"], periodic=[0] * D)
    sp.finish_implicit_step((2,) * D, (2,) * D)
    a = sp.download(); sp.destroy(); grid.destroy()
    grid, sp = make_gpu(pgpu, prob, INTERPS["CIC"], periodic=[0] * D)
    sp.advance_velocities_2nd_half(); sp.advance_positions_2nd_half(); sp.apply_bcs((2,) * D, (2,) * D)
    b = sp.download(); sp.destroy(); grid.destroy()
    for name in ("x", "xold", "v", "vold"):
        assert np.array_equal(a[name], b[name]), name


@pytest.mark.parametrize("user", ["download", "advance", "advance_twice", "advance_pos_only", "sort_advance",
                                  "second_half", "collide", "generic_kernel", "ragged"])
def test_update_old_without_copy(pgpu, user):
    """updateOldParticlePositions/Velocities (PicChargedSpecies.cpp:1821-1867) only record "old == new";
    the CC1 tile kernel consumes the alias (out-of-place write + pointer swap), every other user gets
    the copy first.  Whatever runs next must behave as if the copy had been made."""
    n = 5000 if user != "ragged" else 512 * 3 + 77
    prob = _prob(seed=41, n=n)
    _inside(prob)
    interp = INTERPS["CC1"] if user != "generic_kernel" else INTERPS["CC0"]
    grid, sp = make_gpu(pgpu, prob, interp, fnorm=-0.7, cvac_norm=0.9986)
    if user == "sort_advance":
        sp.bin_particles()
    if user != "advance_pos_only":
        sp.update_old_velocities()
    sp.update_old_positions()
    xold = prob.x.copy()
    vold = prob.v.copy() if user != "advance_pos_only" else prob.vold.copy()
    if user == "sort_advance":
        sp.bin_particles()                       # alias survives a sort
    if user == "download":
        got = sp.download(); o = got["id"].astype(np.int64)
        assert np.array_equal(got["xold"], xold[:, o]) and np.array_equal(got["vold"], vold[:, o])
        assert np.array_equal(got["x"], prob.x[:, o]) and np.array_equal(got["v"], prob.v[:, o])
    elif user == "second_half":
        sp.advance_positions_2nd_half(); sp.advance_velocities_2nd_half()
        got = sp.download()
        assert np.array_equal(got["x"], prob.x) and np.array_equal(got["v"], prob.v)   # 2x - x
        assert np.array_equal(got["xold"], xold) and np.array_equal(got["vold"], vold)
    elif user == "collide":
        sp.bin_particles(); sp.set_moments()
        pgpu.collide_ta(sp, sp, 3.0, 1e-18, 7, 0)
        got = sp.download(); o = got["id"].astype(np.int64)
        assert np.array_equal(got["vold"], vold[:, o])           # the copy happened before v changed
        assert not np.array_equal(got["v"], prob.v[:, o])
    else:
        ie = orc.CC1 if user != "generic_kernel" else orc.CC0
        reps = 2 if user == "advance_twice" else 1
        x, v = prob.x.copy(), prob.v.copy()
        for _ in range(reps):
            sp.advance_iteratively(0.5, deposit=True)
            rc, _, _, _ = orc.advance_particles_iteratively(prob.geom, ie, x, xold, v, vold, prob.E, prob.B,
                                                            -0.7, 0.5 * 0.9986, 1e-12, 21)
            assert rc == 0
        J = [sp.current_get(c) for c in range(3)]
        J0 = prob.new_J()
        orc.deposit_current(prob.geom, ie, x, xold, v, prob.w, 0.5 * 0.9986, J0)
        got = sp.download(); o = got["id"].astype(np.int64)
        assert np.array_equal(got["xold"], xold[:, o]) and np.array_equal(got["vold"], vold[:, o])
        assert np.max(np.abs(got["x"] - x[:, o]) / np.array(prob.dx)[:, None]) <= 4e-12
        assert np.max(np.abs(got["v"] - v[:, o])) / np.max(np.abs(v)) <= 1e-11
        for c in range(3):
            ref = J0[c].a * (-1.0)               # charge / volume_scale
            assert np.max(np.abs(J[c] - ref)) <= 1e-11 * np.max(np.abs(ref))
    sp.destroy(); grid.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("D", [1, 2])
def test_linear_record_adapter_matches_reference_wire_format(pgpu, D):
    """pgpu_species_download_linear / upload_linear speak JustinsParticle::linearOut / linearIn: the field order is the
    one pinned on the reference's own compiled JustinsParticle.cpp (tests/golden/ref_pins.npz: out_wire, 2D), and the
    round trip through the records is the identity."""
    import os
    rng = np.random.default_rng(5)
    ncell = (16,) * D
    grid = pgpu.Grid(D, ncell, (0.0,) * D, (0.25,) * D, 2, (1,) * D)
    sp = pgpu.Species(grid, 1.0, -1.0, 1.0, 1.0)
    n = 1000
    x = rng.random((D, n)) * 4.0
    xold = rng.random((D, n)) * 4.0
    v, vold, w = rng.standard_normal((3, n)), rng.standard_normal((3, n)), rng.random(n) + 0.5
    ids = rng.integers(1, 2 ** 40, n).astype(np.uint64)
    if D == 2:
        gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_pins.npz"))["out_wire"]
        # the particle the reference vector was made from (make_ref_golden.py)
        w[0], x[:, 0], xold[:, 0] = gold[0], gold[1:3], gold[3:5]
        v[:, 0], vold[:, 0], ids[0] = gold[7:10], gold[10:13], np.uint64(gold[13])
    sp.upload(x, v, w, xold=xold, vold=vold, ids=ids)
    rec = sp.download_linear()
    assert rec.shape == (n, 2 * D + 10)
    want = np.concatenate([w[None], x, xold, np.zeros((2, n)), v, vold, ids.astype(np.float64)[None]]).T
    assert np.array_equal(rec, want)
    if D == 2:
        assert np.array_equal(rec[0], gold)                      # bit for bit the reference's record
    sp2 = pgpu.Species(grid, 1.0, -1.0, 1.0, 1.0)
    sp2.upload_linear(rec[::-1])
    got = sp2.download()
    assert sp2.n == n
    for k, a in (("x", x), ("xold", xold), ("v", v), ("vold", vold)):
        assert np.array_equal(got[k], a[:, ::-1]), k
    assert np.array_equal(got["w"], w[::-1]) and np.array_equal(got["id"], ids[::-1])
    sp2.upload_linear(np.zeros((0, 2 * D + 10)))
    assert sp2.n == 0
    sp.destroy(); sp2.destroy(); grid.destroy()
