"""CPU tests that pin the oracle's mass-matrix restatement (oracle/oracle_massmatrix.cpp).

The reference ships no golden vectors for this path, so the restatement is checked against what the mass
matrices are *for* (PicSpeciesInterface::computeJfromMassMatrices, PicSpeciesInterface.cpp:567-753): with the
particle orbits frozen, the Boris half step is affine in E, hence

    deposit_CC1( rho_p * Boris(u_old, E_p(E), B_p) )  ==  J0 + sigma (E - E0)        for every E,

where J0 = deposit_CC1(rho_p * ubar(E0)).  The left side only uses the already tested gather / Boris / CC1
deposit of the oracle; the right side uses cc1_{1,2}d_deposit_mass_matrix, compute_mm_kernals and
compute_J{x,y,z}_from_mass_matrix.  Any wrong weight, component index Nc, stagger shift or contraction offset
breaks the identity at O(1).
"""
import numpy as np
import pytest

from common import Problem, orc

CC1 = orc.CC1


def _setup(D, seed, max_disp, nghost, n=400, ncell=None):
    ncell = ncell or ((12,) if D == 1 else (9, 7))
    dx = (0.25,) if D == 1 else (0.25, 0.2)
    xmin = (-0.5,) if D == 1 else (-0.5, 0.3)
    prob = Problem(D, ncell, dx, xmin, nghost, n, seed=seed, max_disp=max_disp, E0=1.0, B0=2.0)
    return prob


def _ubar(prob, E, fnorm, cnormDt):
    rc, Ep, Bp = orc.gather(prob.geom, CC1, prob.x, prob.xold, E, prob.B)
    assert rc == 0
    return orc.boris(prob.vold.copy(), prob.vold, Ep, Bp, fnorm, cnormDt, True), Bp


def _identity(D, seed, max_disp, nghost):
    prob = _setup(D, seed, max_disp, nghost)
    fnorm, cnormDt, qovs = -0.8, 0.3, -1.7
    alphas = fnorm * cnormDt / 2.0
    # linearisation point
    ubar0, _ = _ubar(prob, prob.E, fnorm, cnormDt)
    nc, sigma = orc.mm_alloc(D, CC1, nghost, prob.box_lo, prob.box_hi)
    J0 = prob.new_J()
    rc = orc.deposit_mass_matrices(prob.geom, CC1, prob.x, prob.xold, ubar0, prob.vold, prob.w, qovs, alphas,
                                   cnormDt, prob.B, J0, sigma)
    assert rc == 0
    # J0 is the ordinary CC1 current deposit of w*qovs*ubar0
    Jd = prob.new_J()
    assert orc.deposit_current(prob.geom, CC1, prob.x, prob.xold, ubar0, prob.w * qovs, cnormDt, Jd) == 0
    for c in range(3):
        s = np.max(np.abs(Jd[c].a))
        assert np.max(np.abs(J0[c].a - Jd[c].a)) < 2e-13 * s, c
    # perturbed field (not periodic on purpose: ghosts are independent unknowns of the contraction)
    rng = np.random.default_rng(seed + 100)
    E1 = [f.copy() for f in prob.E]
    for f in E1:
        f.a += rng.standard_normal(f.a.shape) * 0.7
    ubar1, _ = _ubar(prob, E1, fnorm, cnormDt)
    Jdirect = prob.new_J()
    assert orc.deposit_current(prob.geom, CC1, prob.x, prob.xold, ubar1, prob.w * qovs, cnormDt, Jdirect) == 0
    Jmm = prob.new_J()
    orc.compute_J_from_mass_matrices(D, nc, sigma, prob.E, E1, J0, Jmm)
    for c in range(3):
        s = np.max(np.abs(Jdirect[c].a))
        err = np.max(np.abs(Jmm[c].a - Jdirect[c].a))
        assert err < 5e-13 * s, (c, err, s)
        # and the perturbation is not trivially small
        assert np.max(np.abs(Jdirect[c].a - J0[c].a)) > 1e-3 * s
    return prob, nc, sigma


@pytest.mark.parametrize("max_disp,nghost", [(0.4, 2), (0.95, 2), (1.9, 3)])
def test_cc1_1d_mass_matrix_reproduces_perturbed_current(max_disp, nghost):
    _identity(1, 11, max_disp, nghost)


@pytest.mark.parametrize("max_disp,nghost", [(0.4, 3), (0.95, 3), (1.9, 4)])
def test_cc1_2d_mass_matrix_reproduces_perturbed_current(max_disp, nghost):
    _identity(2, 21, max_disp, nghost)


def test_ncomp_follows_the_reference_tables():
    # PicSpeciesInterface.cpp:311-318 (1D) and :320-349 (2D), ghosts 3 -> maxXings 2 (1D) / 1 (2D)
    nc = orc.mm_ncomp(1, CC1, 3)
    assert nc[:, 0].tolist() == [7, 6, 6, 6, 3, 3, 6, 3, 3]
    nc = orc.mm_ncomp(2, CC1, 3)
    assert nc.tolist() == [[5, 7], [6, 6], [4, 5], [6, 6], [7, 5], [5, 4], [4, 5], [5, 4], [3, 3]]
    with pytest.raises(ValueError):
        orc.mm_ncomp(2, CC1, 2)


def test_too_many_crossings_is_an_error():
    prob = _setup(1, 5, 7.0, 2)
    nc, sigma = orc.mm_alloc(1, CC1, 2, prob.box_lo, prob.box_hi)
    J0 = prob.new_J()
    rc = orc.deposit_mass_matrices(prob.geom, CC1, prob.x, prob.xold, prob.v, prob.vold, prob.w, 1.0, 0.1, 0.3,
                                   prob.B, J0, sigma)
    assert rc == -1


def test_mm_kernels_are_the_boris_response():
    """alphas * d(ubar)/dE of PicSpeciesUtils::applyForces equals f / rhop (non-relativistic)."""
    rng = np.random.default_rng(3)
    fnorm, cnormDt = 1.3, 0.4
    alphas = fnorm * cnormDt / 2
    for _ in range(20):
        Bp = rng.standard_normal(3) * 2
        uo = rng.standard_normal(3) * 0.1
        Ep = rng.standard_normal(3)
        ub = orc.boris(uo[:, None].copy(), uo[:, None], Ep[:, None], Bp[:, None], fnorm, cnormDt, True)[:, 0]
        qp, vol = 0.37, 0.05
        fp, f = orc.mm_kernels(Bp, qp, alphas, vol, uo, ub)
        rhop = qp / vol
        assert np.allclose(fp, rhop * ub, rtol=1e-15)
        for e in range(3):
            dE = np.zeros(3)
            dE[e] = 1.0
            ub2 = orc.boris(uo[:, None].copy(), uo[:, None], (Ep + dE)[:, None], Bp[:, None], fnorm, cnormDt, True)[:, 0]
            assert np.allclose(rhop * (ub2 - ub), f[:, e], rtol=0, atol=1e-13 * rhop)


def test_relativistic_kernels_reduce_to_the_classical_ones_at_low_energy():
    rng = np.random.default_rng(4)
    Bp = rng.standard_normal(3)
    uo = rng.standard_normal(3) * 1e-3
    ub = uo + rng.standard_normal(3) * 1e-4
    fp0, f0 = orc.mm_kernels(Bp, 0.3, 0.1, 0.05, uo, ub, relativistic=False)
    fp1, f1 = orc.mm_kernels(Bp, 0.3, 0.1, 0.05, uo, ub, relativistic=True)
    assert np.allclose(fp0, fp1, rtol=1e-5)
    assert np.allclose(f0, f1, rtol=1e-5, atol=1e-8 * np.max(np.abs(f0)))
    # above the gamma threshold of the reference (1.01) the correction terms switch on
    uo = np.array([0.5, -0.3, 0.2])
    ub = np.array([0.55, -0.25, 0.22])
    _, f0 = orc.mm_kernels(Bp, 0.3, 0.1, 0.05, uo, ub, relativistic=False)
    _, f1 = orc.mm_kernels(Bp, 0.3, 0.1, 0.05, uo, ub, relativistic=True)
    assert np.max(np.abs(f1 - f0)) > 1e-3 * np.max(np.abs(f0))
