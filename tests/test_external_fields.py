"""addExternalFieldsToParticles (PicChargedSpecies.cpp:3948-3996): the six external-field grid functions
of EMFields::getExternalE/B (EMFields.H:176-198) -- Constant, Cosine, Heavyside (ibc/grid_functions) -- added to E_p, B_p
after every gather of the particle loop (:1606, :1652, :1669)."""
import numpy as np
import pytest

from common import orc, Problem, make_gpu, rel_err, INTERPS


def make_problem(D, seed, max_disp=0.3):
    if D == 1:
        return Problem(1, (24,), (0.25,), (0.5,), 4, 3000, seed=seed, max_disp=max_disp, E0=0.3, B0=0.8)
    return Problem(2, (12, 10), (0.25, 0.3), (0.5, -1.0), 4, 3000, seed=seed, max_disp=max_disp, E0=0.3, B0=0.8)

SIX = [
    {"type": 2, "value": 1.7, "constant": 0.3, "L": (3.0, 2.0), "mode": (1.0, 2.0), "phase": (0.25, 0.5)},    # Ex: Cosine
    {"type": 1, "value": -0.8},                                                                                 # Ey: Constant
    {"type": 0},                                                                                                # Ez: none
    {"type": 3, "C": (0.5, 1.0), "A": (2.0, -0.5), "X0": (1.1, 0.9), "eps": (1e-9, 1e-9)},                      # Bx: Heavyside
    {"type": 2, "value": 0.9, "constant": -0.1, "L": (1.5, 2.5), "mode": (2.0, 1.0), "phase": (0.0, 1.0)},      # By: Cosine
    {"type": 1, "value": 2.2},                                                                                  # Bz: Constant
]


def _numpy_value(d, D, x):
    t = d.get("type", 0)
    if t == 0:
        return np.zeros(x.shape[1])
    if t == 1:
        return np.full(x.shape[1], d["value"])
    if t == 2:
        v = np.full(x.shape[1], d["value"])
        for k in range(D):
            arg = np.fmod(2 * np.pi * d["mode"][k] * x[k] / d["L"][k] + d["phase"][k] * np.pi, 2 * np.pi)
            v = v * np.cos(arg)
        return v + d["constant"]
    v = np.ones(x.shape[1])
    for k in range(D):
        arg = x[k] - d["X0"][k]
        H = np.where(arg < 0.0, 0.0, 1.0)
        H = np.where(np.abs(arg) < d["eps"][k], 0.5, H)
        v = v * (d["C"][k] + d["A"][k] * H)
    return v


@pytest.mark.parametrize("D", [1, 2])
def test_oracle_external_field_values(D):
    """The oracle's grid functions against their closed forms; off by default and after clearing."""
    prob = make_problem(D, seed=5)
    Ep, Bp = np.zeros((3, prob.n)), np.zeros((3, prob.n))
    orc.add_external_fields(D, prob.x, Ep, Bp)
    assert not Ep.any() and not Bp.any()
    orc.set_external_fields(SIX)
    try:
        orc.add_external_fields(D, prob.x, Ep, Bp)
    finally:
        orc.set_external_fields(None)
    for c in range(3):
        assert np.allclose(Ep[c], _numpy_value(SIX[c], D, prob.x), rtol=0, atol=1e-15)
        assert np.allclose(Bp[c], _numpy_value(SIX[3 + c], D, prob.x), rtol=0, atol=1e-15)
    assert np.ptp(Ep[0]) > 0.5 and np.ptp(Bp[0]) > 0.1


def _ext_structs(pgpu):
    out = []
    for d in SIX:
        f = pgpu.ExtFn()
        f.type = d.get("type", 0)
        f.value, f.constant = d.get("value", 0.0), d.get("constant", 0.0)
        for k in ("L", "mode", "phase", "C", "A", "X0", "eps"):
            v = d.get(k, (0.0, 0.0))
            getattr(f, k)[0], getattr(f, k)[1] = v[0], v[1]
        out.append(f)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("D", [1, 2])
def test_gpu_add_external_fields_stored(pgpu, D):
    prob = make_problem(D, seed=6)
    grid, sp = make_gpu(pgpu, prob, INTERPS["CIC"])
    grid.set_external_fields(_ext_structs(pgpu))
    sp.interpolate_fields()
    E0, B0 = sp.particle_fields()
    sp.add_external_fields()
    E1, B1 = sp.particle_fields()
    for c in range(3):
        assert np.abs((E1[c] - E0[c]) - _numpy_value(SIX[c], D, prob.x)).max() < 1e-14
        assert np.abs((B1[c] - B0[c]) - _numpy_value(SIX[3 + c], D, prob.x)).max() < 1e-14
    sp.destroy(); grid.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("iterative", [True, False])
@pytest.mark.parametrize("interp", ["CIC", "CC1"])
@pytest.mark.parametrize("D", [1, 2])
def test_gpu_advance_with_external_fields_matches_oracle(pgpu, D, interp, iterative):
    """The particle loop with external fields on: CUDA against the oracle, and the external fields do change the result."""
    prob = make_problem(D, seed=7)
    fn, dt, cv = -0.1, 0.5, 0.9986
    grid, sp = make_gpu(pgpu, prob, INTERPS[interp], fnorm=fn, cvac_norm=cv, iter_max=(21 if iterative else 0))
    sp.advance_iteratively(dt, deposit=False) if iterative else sp.advance_particles(dt)
    plain = sp.download()
    sp.destroy()
    grid2, sp2 = make_gpu(pgpu, prob, INTERPS[interp], fnorm=fn, cvac_norm=cv, iter_max=(21 if iterative else 0))
    grid2.set_external_fields(_ext_structs(pgpu))
    st = sp2.advance_iteratively(dt, deposit=False) if iterative else sp2.advance_particles(dt)
    got = sp2.download()
    x, v = prob.x.copy(), prob.v.copy()
    orc.set_external_fields(SIX)
    try:
        if iterative:
            rc, _, unconv, its = orc.advance_particles_iteratively(prob.geom, INTERPS[interp], x, prob.xold, v, prob.vold,
                                                                   prob.E, prob.B, fn, dt * cv, 1e-12, 21)
            # a particle whose orbit straddles the jump of the Heavyside function can cycle for ever: both sides must
            # agree on that, and it is left out of the value comparison
            assert unconv <= 3 and st.num_unconverged == unconv
            ok = its < 22
        else:
            rc = orc.advance_particles(prob.geom, INTERPS[interp], x, prob.xold, v, prob.vold, prob.E, prob.B, fn,
                                       dt * cv, 0)
    finally:
        orc.set_external_fields(None)
    assert rc == 0
    if not iterative:
        ok = np.ones(prob.n, dtype=bool)
    assert rel_err(got["v"][:, ok], v[:, ok]) < 1e-11 and np.abs(got["x"] - x)[:, ok].max() < 4e-12 * prob.dx[0]
    assert rel_err(plain["v"], v) > 1e-3      # the external fields matter
    sp2.destroy(); grid.destroy(); grid2.destroy()
