"""GPU parity of the mass-matrix path against the CPU oracle, through the C ABI.

pgpu_accumulate_mass_matrices  <->  orc.deposit_mass_matrices   (cc1_{1,2}d_deposit_mass_matrix + compute_mm_kernals)
pgpu_compute_J_from_mass_matrices <-> orc.compute_J_from_mass_matrices (compute_J{x,y,z}_from_mass_matrix)

The CUDA file is built without FMA contraction, so every per-particle product equals the oracle's; the sums differ
by the order of the atomics only.  Bar: 1e-12 of the array's scale (the north star's push/deposit tolerance),
component indices exact (a wrong Nc would show up at O(1)).
"""
import numpy as np
import pytest

from common import Problem, make_gpu, orc, rel_err

pytestmark = pytest.mark.gpu
CC1 = orc.CC1
TOL = 1e-12


def _prob(D, seed, n, max_disp, nghost):
    if D == 1:
        return Problem(1, (24,), (0.25,), (0.5,), nghost, n, seed=seed, max_disp=max_disp, B0=2.0)
    return Problem(2, (12, 10), (0.25, 0.3), (0.5, -1.0), nghost, n, seed=seed, max_disp=max_disp, B0=2.0)


def _sorted(prob):
    """Cell-sort the particles the way pgpu_bin_particles would (by dual cell), so that the run kernel sees runs."""
    D = prob.D
    key = np.zeros(prob.n, dtype=np.int64)
    for d in range(D):
        key = key * 4096 + np.floor((prob.x[d] - prob.xmin[d]) / prob.dx[d] * 2).astype(np.int64)
    o = np.argsort(key, kind="stable")
    for name in ("x", "xold", "v", "vold"):
        setattr(prob, name, np.ascontiguousarray(getattr(prob, name)[:, o]))
    prob.w = np.ascontiguousarray(prob.w[o])
    return prob


def _oracle(prob, nghost, charge, vs, fnorm, cvac, dt, relativistic=False):
    nc, sigma = orc.mm_alloc(prob.D, CC1, nghost, prob.box_lo, prob.box_hi)
    J0 = prob.new_J()
    cnormDt = dt * cvac
    rc = orc.deposit_mass_matrices(prob.geom, CC1, prob.x, prob.xold, prob.v, prob.vold, prob.w, charge / vs,
                                   fnorm * cnormDt / 2.0, cnormDt, prob.B, J0, sigma, relativistic=relativistic)
    assert rc == 0
    return nc, sigma, J0


def _check(grid, nc, sigma, J0):
    for c in range(3):
        assert rel_err(grid.mass_matrix_J0_get(c), J0[c].a) <= TOL, ("J0", c)
    for k in range(9):
        got = grid.mass_matrix_get(k)
        ref = sigma[k].a
        assert got.shape == ref.shape
        assert np.max(np.abs(ref)) > 0
        assert rel_err(got, ref) <= TOL, orc.SIGMA_NAMES[k]


@pytest.mark.parametrize("mode", ["run", "generic"])
@pytest.mark.parametrize("D,max_disp,nghost", [(1, 0.4, 2), (1, 1.9, 3), (2, 0.4, 3), (2, 0.95, 3), (2, 1.9, 4)])
def test_accumulate_mass_matrices(pgpu, D, max_disp, nghost, mode):
    prob = _sorted(_prob(D, 31, 6000, max_disp, nghost))
    charge, vs, fnorm, cvac, dt = -1.0, 2.5, 0.8, 1.3, 0.2
    pgpu.load().pgpu_set_deposit_mode(1 if mode == "run" else 0)
    grid, sp = make_gpu(pgpu, prob, CC1, charge=charge, volume_scale=vs, fnorm=fnorm, cvac_norm=cvac)
    try:
        nc_gpu = grid.mass_matrices_init(CC1)
        nc, sigma, J0 = _oracle(prob, nghost, charge, vs, fnorm, cvac, dt)
        assert nc_gpu.tolist() == nc.tolist()
        grid.mass_matrices_zero()
        sp.accumulate_mass_matrices(dt)
        _check(grid, nc, sigma, J0)
        # accumulate again (second species of the same kind): everything doubles
        sp.accumulate_mass_matrices(dt)
        got = grid.mass_matrix_get(0)
        assert rel_err(got, 2.0 * sigma[0].a) <= TOL
    finally:
        pgpu.load().pgpu_set_deposit_mode(1)
        sp.destroy(); grid.destroy()


@pytest.mark.parametrize("D", [1, 2])
@pytest.mark.parametrize("chunk", ["3", "32"])
def test_runs_carried_across_tiles(pgpu, monkeypatch, chunk, D):
    """Large problems give a warp several consecutive 32-particle tiles and keep a run open across them (and across
    deferred particles); PGPU_MM_CHUNK forces that path on a problem the oracle can follow."""
    monkeypatch.setenv("PGPU_MM_CHUNK", chunk)
    prob = _sorted(_prob(D, 37, 9001, 0.95, 3))
    grid, sp = make_gpu(pgpu, prob, CC1, charge=-1.0, fnorm=0.9)
    try:
        grid.mass_matrices_init(CC1)
        nc, sigma, J0 = _oracle(prob, 3, -1.0, 1.0, 0.9, 1.0, 0.15)
        sp.accumulate_mass_matrices(0.15)
        _check(grid, nc, sigma, J0)
    finally:
        sp.destroy(); grid.destroy()


def test_unsorted_particles_and_ragged_count(pgpu):
    """Any particle order is correct (runs of length one), also when n is not a multiple of the warp size."""
    prob = _prob(2, 32, 1237, 0.6, 3)
    grid, sp = make_gpu(pgpu, prob, CC1, charge=1.0, fnorm=-0.5)
    try:
        grid.mass_matrices_init(CC1)
        nc, sigma, J0 = _oracle(prob, 3, 1.0, 1.0, -0.5, 1.0, 0.3)
        sp.accumulate_mass_matrices(0.3)
        _check(grid, nc, sigma, J0)
    finally:
        sp.destroy(); grid.destroy()


def test_J_from_mass_matrices_matches_oracle_and_the_perturbed_deposit(pgpu):
    """computeJfromMassMatrices on the device; and the identity the matrices exist for: with frozen orbits,
    J0 + sigma (E - E0) equals the CC1 deposit of the Boris response to E."""
    nghost = 3
    prob = _sorted(_prob(2, 33, 5000, 0.95, nghost))
    charge, vs, fnorm, cvac, dt = -1.0, 1.0, 0.8, 1.0, 0.3
    cnormDt = dt * cvac
    # linearisation point: ubar = Boris(E0) at the stored orbits
    rc, Ep, Bp = orc.gather(prob.geom, CC1, prob.x, prob.xold, prob.E, prob.B)
    assert rc == 0
    prob.v = orc.boris(prob.vold.copy(), prob.vold, Ep, Bp, fnorm, cnormDt, True)
    grid, sp = make_gpu(pgpu, prob, CC1, charge=charge, volume_scale=vs, fnorm=fnorm, cvac_norm=cvac)
    try:
        grid.mass_matrices_init(CC1)
        grid.mass_matrices_zero()
        sp.accumulate_mass_matrices(dt)
        grid.mass_matrices_save_E0()
        nc, sigma, J0 = _oracle(prob, nghost, charge, vs, fnorm, cvac, dt)
        rng = np.random.default_rng(5)
        E1 = [f.copy() for f in prob.E]
        for f in E1:
            f.a += rng.standard_normal(f.a.shape) * 0.5
        grid.fields_select(1)
        grid.set_fields([(f.lo, f.hi, f.a) for f in E1], [(f.lo, f.hi, f.a) for f in prob.B])
        grid.compute_J_from_mass_matrices()
        Jref = prob.new_J()
        orc.compute_J_from_mass_matrices(2, nc, sigma, prob.E, E1, J0, Jref)
        # direct: deposit of the Boris response to E1
        rc, Ep1, Bp1 = orc.gather(prob.geom, CC1, prob.x, prob.xold, E1, prob.B)
        ub1 = orc.boris(prob.vold.copy(), prob.vold, Ep1, Bp1, fnorm, cnormDt, True)
        Jd = prob.new_J()
        assert orc.deposit_current(prob.geom, CC1, prob.x, prob.xold, ub1, prob.w * charge / vs, cnormDt, Jd) == 0
        for c in range(3):
            J = grid.current_get(c)
            assert rel_err(J, Jref[c].a) <= TOL, c
            assert rel_err(J, Jd[c].a) <= 5e-12, c
        grid.fields_select(0)
    finally:
        sp.destroy(); grid.destroy()


def test_relativistic_species(pgpu):
    prob = _sorted(_prob(2, 34, 3000, 0.6, 3))
    prob.vold *= 8.0          # gamma up to ~1.3: the correction terms above 1.01 are exercised
    prob.v = prob.vold + 0.05 * prob.v
    grid = pgpu.Grid(2, prob.ncell, prob.xmin, prob.dx, 3, [1, 1])
    E, B = prob.fields_for_gpu()
    grid.set_fields(E, B)
    sp = pgpu.Species(grid, 1.0, -1.0, 0.7, 1.0, interp_N=1, interp_J=CC1, interp_E=CC1, relativistic=1)
    sp.upload(prob.x, prob.v, prob.w, xold=prob.xold, vold=prob.vold, ids=np.arange(prob.n, dtype=np.uint64))
    try:
        grid.mass_matrices_init(CC1)
        sp.accumulate_mass_matrices(0.25)
        nc, sigma, J0 = _oracle(prob, 3, -1.0, 1.0, 0.7, 1.0, 0.25, relativistic=True)
        _check(grid, nc, sigma, J0)
    finally:
        sp.destroy(); grid.destroy()


def test_too_many_crossings_is_reported(pgpu):
    prob = _prob(2, 35, 500, 3.9, 3)   # up to 3 faces per direction, maxXings = 1
    grid, sp = make_gpu(pgpu, prob, CC1)
    try:
        grid.mass_matrices_init(CC1)
        sp.accumulate_mass_matrices(0.1)
        with pytest.raises(pgpu.PgpuError) as e:
            grid.mass_matrix_get(0)
        assert e.value.code == -3   # PGPU_ERR_SEGMENTS
    finally:
        pgpu.load().pgpu_picard_totals(None, None, None, 1)
        sp.destroy(); grid.destroy()


def test_needs_init_and_cc1(pgpu):
    prob = _prob(1, 36, 100, 0.3, 2)
    grid, sp = make_gpu(pgpu, prob, CC1)
    try:
        with pytest.raises(pgpu.PgpuError):
            sp.accumulate_mass_matrices(0.1)
        with pytest.raises(pgpu.PgpuError):
            grid.mass_matrices_init(orc.TSC)
    finally:
        sp.destroy(); grid.destroy()


def test_two_boxes_J_from_mass_matrices_matches_single_box(pgpu):
    """SURVEY 8(e) for the mass-matrix path: each box accumulates the matrices of its own particles and contracts them
    over its ghosted box; the ghost add-exchange of J (peer-memory halo, as after a deposit) then gives the single-box
    current.  The sigmas themselves are never exchanged -- as in PicSpeciesInterface::setMassMatrices."""
    from picnic_b200 import halo
    ncell, ng, dx, xmin = (32, 16), 3, (0.25, 0.5), (0.0, -1.0)
    prob = Problem(2, ncell, dx, xmin, ng, 20000, seed=41, max_disp=0.8, B0=2.0)
    rng = np.random.default_rng(42)
    E1 = [f.copy() for f in prob.E]
    for f, st in zip(E1, orc.E_STAG[2]):       # a periodic perturbation, so that ghosts stay images
        core = rng.standard_normal(ncell) * 0.5
        idx = [np.mod(np.arange(f.lo[d], f.hi[d] + 1), ncell[d]) for d in range(2)]
        f.a += core[np.ix_(idx[0], idx[1])]
    dt, charge, fnorm = 0.2, -1.0, 0.8

    def sub(fabs, g):
        out = []
        for c, f in enumerate(fabs):
            lo, hi = g.field_bounds(c if fabs is not prob.B else 3 + c)
            ii = np.mod(np.arange(lo[0], hi[0] + 1), ncell[0]) - f.lo[0]
            jj = np.mod(np.arange(lo[1], hi[1] + 1), ncell[1]) - f.lo[1]
            out.append((lo, hi, np.asfortranarray(f.a[np.ix_(ii, jj)])))
        return out

    def run(g, m):
        g.set_fields(sub(prob.E, g), sub(prob.B, g))
        sp = pgpu.Species(g, 1.0, charge, fnorm, 1.0, interp_N=1, interp_J=CC1, interp_E=CC1)
        sp.upload(prob.x[:, m], prob.v[:, m], prob.w[m], xold=prob.xold[:, m], vold=prob.vold[:, m],
                  ids=np.arange(prob.n, dtype=np.uint64)[m])
        g.mass_matrices_init(CC1)
        g.mass_matrices_zero()
        sp.accumulate_mass_matrices(dt)
        g.mass_matrices_save_E0()
        g.fields_select(1)
        g.set_fields(sub(E1, g), sub(prob.B, g))
        g.compute_J_from_mass_matrices()
        return sp

    g1 = pgpu.Grid(2, ncell, xmin, dx, ng, (1, 1))
    s1 = run(g1, np.ones(prob.n, dtype=bool))
    g1.current_finalize()
    Jg = [(g1.field_bounds(c), g1.current_get(c)) for c in range(3)]
    s1.destroy(); g1.destroy()

    lay = halo.BoxLayout(2, ncell, (16, 16), ng, (1, 1))
    own = np.floor((prob.xold[0] - xmin[0]) / (dx[0] * 16)).astype(int)
    grids, sps, hxs = [], [], []
    for r in range(lay.world):
        lo, hi = lay.box(r)
        g = pgpu.Grid(2, ncell, xmin, dx, ng, (1, 1), box_lo=lo, box_hi=hi)
        sps.append(run(g, own == r))
        grids.append(g)
        hxs.append(halo.PeerHaloExchange(lay, r, g))
    halo.PeerHaloExchange.connect_local(hxs)
    for h in hxs:
        h.begin()
    for ph in range(hxs[0].nphase):
        for h in hxs:
            h.send(ph)
        for h in hxs:
            h.recv_add(ph)
    worst = 0.0
    for g in grids:
        g.current_finalize()
        for c in range(3):
            lo, hi = g.field_bounds(c)
            a = g.current_get(c)
            (glo, _), ga = Jg[c]
            ii = np.mod(np.arange(lo[0], hi[0] + 1), ncell[0]) - glo[0]
            jj = np.mod(np.arange(lo[1], hi[1] + 1), ncell[1]) - glo[1]
            worst = max(worst, float(np.max(np.abs(a - ga[np.ix_(ii, jj)])) / np.max(np.abs(ga))))
    for h in hxs:
        h.destroy()
    for s in sps:
        s.destroy()
    for g in grids:
        g.destroy()
    assert worst < 1e-12, worst
