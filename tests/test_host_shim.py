"""The C++ host side (picnic_b200/host): classes with the reference's method names over the C ABI.
CPU: it builds, links against libpicnic_gpu.so and fails loudly without a device (the reference's
MayDay::Error convention: message + non-zero exit).  GPU: picnic_b200/host/example_driver.cpp runs
the particle side of one theta-implicit step through those classes; J and the particles must match
the oracle driven through the same call sequence."""
import os
import struct
import subprocess

import numpy as np
import pytest

from common import orc, decks, ROOT

HOST = os.path.join(ROOT, "picnic_b200", "host")


@pytest.fixture(scope="module")
def driver():
    from picnic_b200 import build
    build.build()
    subprocess.check_call(["make", "-s", "-C", HOST])
    return os.path.join(HOST, "example_driver")


def test_host_shim_builds_and_fails_loudly_without_device(driver, tmp_path):
    import torch
    assert os.path.exists(os.path.join(HOST, "libpicgpuhost.a"))
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = tmp_path / "p.bin"
    p.write_bytes(struct.pack("8i", 2, 8, 8, 2, 0, 1, 0, 0) + struct.pack("10d", 0, 0, 1, 1, 0.1, 1e-12, 1, 1, 1, 0))
    r = subprocess.run([driver, str(p), str(tmp_path / "r.bin")], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_example_driver_matches_oracle(driver, tmp_path):
    deck = decks.deck_c3(ncell=16, ppc=3, dt=2.0)
    lo, hi = (0, 0), (15, 15)
    E, B = decks.analytic_fields(deck, lo, hi)
    sdef = deck.species[0]
    rng = np.random.default_rng(5)
    p = decks.load_species(deck, sdef, lo, hi, rng)
    n, n_outer = p["w"].size, 2
    fn = sdef.fnorm_const(deck.units)
    prob = tmp_path / "p.bin"
    with open(prob, "wb") as f:
        f.write(struct.pack("8i", 2, 16, 16, deck.nghost, n, n_outer, deck.iter_max, 0))
        f.write(struct.pack("10d", deck.xmin[0], deck.xmin[1], deck.dx[0], deck.dx[1], deck.dt, deck.rtol, fn,
                            deck.units.cvac_norm, deck.volume_scale, 0.0))
        for (_, _, a) in list(E) + list(B):
            f.write(np.asfortranarray(a).tobytes(order="F"))
        f.write(np.ascontiguousarray(p["x"]).tobytes())
        f.write(np.ascontiguousarray(p["v"]).tobytes())
        f.write(np.ascontiguousarray(p["w"]).tobytes())
    res = tmp_path / "r.bin"
    r = subprocess.run([driver, str(prob), str(res)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = np.fromfile(res, dtype=np.float64)
    # Mesh::setMassMatrices / computeJfromMassMatrices at E == E0 give back the deposited current
    mm = [l for l in r.stdout.splitlines() if "mass matrices" in l]
    assert mm and "ncomp_xx=5x7" in mm[0] and float(mm[0].split()[-1]) < 1e-12, r.stdout
    # oracle, same call sequence
    geom = orc.make_geom(2, deck.xmin, deck.xmax, deck.dx, deck.nghost)
    Ef = [orc.Fab(l, h, a) for (l, h, a) in E]
    Bf = [orc.Fab(l, h, a) for (l, h, a) in B]
    x, v = p["x"].copy(), p["v"].copy()
    xold, vold = p["x"].copy(), p["v"].copy()
    for _ in range(n_outer):
        rc, _, unconv, _ = orc.advance_particles_iteratively(geom, deck.interp_E, x, xold, v, vold, Ef, Bf, fn,
                                                             deck.cnorm_dt, deck.rtol, deck.iter_max)
        assert rc == 0
        J0 = [orc.fab_for(lo, hi, deck.nghost, s) for s in orc.E_STAG[2]]
        orc.deposit_current(geom, deck.interp_J, x, xold, v, p["w"], deck.cnorm_dt, J0)
    off = 0
    for c in range(3):
        orc.scale_fab(J0[c], 2, sdef.charge / deck.volume_scale)
        orc.fold_periodic(J0[c], 2, orc.E_STAG[2][c], lo, hi, (1, 1))
        got = raw[off:off + J0[c].a.size].reshape(J0[c].a.shape, order="F")
        off += J0[c].a.size
        assert np.max(np.abs(got - J0[c].a)) <= 1e-11 * np.max(np.abs(J0[c].a)), c
    orc.lib().orc_advance_velocities_2nd_half(n, orc._ptr(v), orc._ptr(vold))
    orc.lib().orc_advance_positions_2nd_half(2, n, orc._ptr(x), orc._ptr(xold))
    L = np.array(deck.xmax) - np.array(deck.xmin)
    xw = np.array(deck.xmin)[:, None] + np.mod(x - np.array(deck.xmin)[:, None], L[:, None])
    gx = raw[off:off + 2 * n].reshape(2, n); off += 2 * n
    gv = raw[off:off + 3 * n].reshape(3, n); off += 3 * n
    ids = raw[off:off + n].view(np.uint64)
    k = np.argsort(ids)                          # binTheParticles reordered the arrays
    assert np.array_equal(ids[k], np.arange(n, dtype=np.uint64))
    assert np.max(np.abs(gx[:, k] - xw)) <= 1e-11 * deck.dx[0]
    assert np.max(np.abs(gv[:, k] - v)) <= 1e-11 * np.max(np.abs(v))
    # the collision tail of the driver: TakizukaAbe::setMeanFreeTime on the end-of-step particles
    line = [l for l in r.stdout.splitlines() if "TA scatterDt" in l][0]
    dt_scatter = float(line.split("scatterDt=")[1].split()[0])
    cells = np.floor((xw - np.array(deck.xmin)[:, None]) / np.array(deck.dx)[:, None]).astype(int)
    cid = cells[0] + 16 * cells[1]
    dV = deck.dx[0] * deck.dx[1] * deck.volume_scale
    dens = np.bincount(cid, weights=p["w"], minlength=256) / dV
    ene = np.stack([np.bincount(cid, weights=0.5 * sdef.mass * p["w"] * v[k_] ** 2, minlength=256) / dV for k_ in range(3)])
    nu = orc.ta_nu_max((dens, np.zeros((3, 256)), ene), (dens, np.zeros((3, 256)), ene), sdef.charge, sdef.charge,
                       sdef.mass, sdef.mass, 3.0, True)
    assert abs(dt_scatter * nu - 1.0) < 1e-9
    assert int(line.split("pairs=")[1]) == sum((c // 2 if c % 2 == 0 else (c - 3) // 2 + 3) for c in np.bincount(cid, minlength=256) if c >= 2)


@pytest.mark.gpu
def test_example_two_boxes(driver):
    """GhostExchange + ParticleMigration of the C++ shim: 2 x 2 boxes in one process against the single box
    (picnic_b200/host/example_two_boxes.cpp checks J everywhere, ownership and the id sum itself)."""
    r = subprocess.run([os.path.join(HOST, "example_two_boxes")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "misplaced 0" in r.stdout and "ids_ok 1" in r.stdout


@pytest.mark.gpu
def test_example_shock_1d_life_cycle_through_the_shim(driver):
    """picnic_b200/host/example_shock_1d.cpp: the implicit shock decks' particle side (inflow list with
    suborbit_inflow_J, outflow list, packed host I/O, List<JustinsParticle> records, the fused explicit step) driven through
    the C++ host classes; the program checks itself and exits non-zero on any failed check."""
    exe = os.path.join(HOST, "example_shock_1d")
    assert os.path.exists(exe)
    r = subprocess.run([exe, "150"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok" in r.stdout and "injected=4500" in r.stdout
