"""ctypes front-end of the CPU oracle (liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module (see oracle/picnic_oracle.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

CIC, TSC, CC0, CC1 = 0, 1, 2, 3

# component centring (1 = nodal, 0 = cell centred) per direction, SURVEY App. A
E_STAG = {1: [(0,), (1,), (1,)], 2: [(0, 1), (1, 0), (1, 1)]}
B_STAG = {1: [(1,), (0,), (0,)], 2: [(1, 0), (0, 1), (0, 0)]}


class Geom(C.Structure):
    _fields_ = [("D", C.c_int), ("le", C.c_double * 2), ("re", C.c_double * 2),
                ("dx", C.c_double * 2), ("ghosts", C.c_int),
                ("bc_lo", C.c_int * 2), ("bc_hi", C.c_int * 2)]


class CFab(C.Structure):
    _fields_ = [("p", C.POINTER(C.c_double)), ("lo", C.c_int * 2), ("hi", C.c_int * 2)]


class CMFab(C.Structure):
    _fields_ = [("p", C.POINTER(C.c_double)), ("lo", C.c_int * 2), ("hi", C.c_int * 2), ("ncomp", C.c_int)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_ta_b90_fact.restype = C.c_double
        dbl = C.c_double
        _LIB.orc_boris.argtypes = [C.c_long] + [C.c_void_p] * 4 + [dbl, dbl, C.c_int]
        _LIB.orc_advance_positions_explicit.argtypes = [C.c_int, C.c_long] + [C.c_void_p] * 3 + [dbl]
        _LIB.orc_advance_positions_implicit.argtypes = [C.c_int, C.c_long] + [C.c_void_p] * 3 + [dbl]
        _LIB.orc_advance_positions_2nd_half.argtypes = [C.c_int, C.c_long, C.c_void_p, C.c_void_p]
        _LIB.orc_advance_velocities_2nd_half.argtypes = [C.c_long, C.c_void_p, C.c_void_p]
        _LIB.orc_average_velocities.argtypes = [C.c_long, C.c_void_p, C.c_void_p]
        _LIB.orc_set_external_fields.argtypes = [C.c_void_p]
        _LIB.orc_gather.argtypes = [C.c_void_p, C.c_int, C.c_long] + [C.c_void_p] * 6
        _LIB.orc_deposit_current.argtypes = [C.c_void_p, C.c_int, C.c_long] + [C.c_void_p] * 4 + [dbl, C.c_void_p]
        _LIB.orc_deposit_rho.argtypes = [C.c_void_p, C.c_int, C.c_long] + [C.c_void_p] * 4
        _LIB.orc_scale_fab.argtypes = [C.c_void_p, C.c_int, dbl]
        _LIB.orc_advance_particles.argtypes = [C.c_void_p, C.c_int, C.c_long] + [C.c_void_p] * 6 + [dbl, dbl, C.c_int]
        _LIB.orc_advance_particles_iteratively.argtypes = (
            [C.c_void_p, C.c_int, C.c_long] + [C.c_void_p] * 6 + [dbl, dbl, dbl, C.c_int] + [C.c_void_p] * 3)
        _LIB.orc_fold_periodic.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
        _LIB.orc_bin.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]
        _LIB.orc_cell_moments.argtypes = [C.c_void_p, C.c_long] + [C.c_void_p] * 3 + [dbl, dbl] + [C.c_void_p] * 5
        _LIB.orc_debye_accumulate.argtypes = [C.c_long] + [C.c_void_p] * 3 + [dbl, dbl, C.c_void_p]
        _LIB.orc_debye_finish.argtypes = [C.c_long, C.c_void_p]
        _LIB.orc_bc_periodic.argtypes = [C.c_long, C.c_void_p, C.c_void_p, dbl, dbl]
        _LIB.orc_bc_symmetry.argtypes = [C.c_long] + [C.c_void_p] * 4 + [dbl, dbl, C.c_int, C.c_int]
        _LIB.orc_rng_seed.argtypes = [C.c_uint64]
        _LIB.orc_scatter_delta_u.argtypes = [dbl] * 7 + [C.c_void_p]
        _LIB.orc_ta_delta_u.argtypes = [C.c_void_p, dbl, C.c_void_p, dbl, dbl, dbl, dbl, dbl, dbl, dbl, C.c_void_p]
        _LIB.orc_ta_b90_fact.argtypes = [dbl] * 4
        _LIB.orc_ta_self.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, dbl, dbl, dbl, dbl, C.c_void_p]
        _LIB.orc_ta_inter.argtypes = ([C.c_long, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, dbl, dbl,
                                       C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, dbl, dbl, dbl, dbl, C.c_void_p])
    return _LIB


def _ptr(a):
    assert a.flags["C_CONTIGUOUS"] or a.flags["F_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class Fab:
    """One component in Chombo FArrayBox layout: column-major, inclusive bounds."""

    def __init__(self, lo, hi, data=None):
        self.lo = tuple(int(v) for v in lo)
        self.hi = tuple(int(v) for v in hi)
        shape = tuple(h - l + 1 for l, h in zip(self.lo, self.hi))
        self.a = np.zeros(shape, dtype=np.float64, order="F") if data is None else np.asfortranarray(data, dtype=np.float64)
        assert self.a.shape == shape

    def c(self):
        f = CFab()
        f.p = self.a.ctypes.data_as(C.POINTER(C.c_double))
        for d in range(2):
            f.lo[d] = self.lo[d] if d < len(self.lo) else 0
            f.hi[d] = self.hi[d] if d < len(self.hi) else 0
        return f

    def copy(self):
        return Fab(self.lo, self.hi, self.a.copy(order="F"))


def fab_for(box_lo, box_hi, nghost, stag):
    lo = [l - nghost for l in box_lo]
    hi = [h + nghost + s for h, s in zip(box_hi, stag)]
    return Fab(lo, hi)


def make_geom(D, le, re, dx, ghosts, bc_lo=None, bc_hi=None):
    g = Geom()
    g.D = D
    for d in range(D):
        g.le[d], g.re[d], g.dx[d] = le[d], re[d], dx[d]
        g.bc_lo[d] = 0 if bc_lo is None else bc_lo[d]
        g.bc_hi[d] = 0 if bc_hi is None else bc_hi[d]
    if D == 1:
        g.le[1], g.re[1], g.dx[1] = 0.0, 1.0, 1.0
    g.ghosts = ghosts
    return g


def _fabs3(fabs):
    arr = (CFab * 3)()
    for i, f in enumerate(fabs):
        arr[i] = f.c()
    return arr


class ExtFn(C.Structure):
    _fields_ = [("type", C.c_int), ("value", C.c_double), ("constant", C.c_double),
                ("L", C.c_double * 2), ("mode", C.c_double * 2), ("phase", C.c_double * 2),
                ("C", C.c_double * 2), ("A", C.c_double * 2), ("X0", C.c_double * 2), ("eps", C.c_double * 2)]


class ExtFields(C.Structure):
    _fields_ = [("on", C.c_int), ("f", ExtFn * 6)]


def set_external_fields(six):
    """six: list of 6 dicts (keys of orc_ext_fn) or None.  Global, like the reference's EMFields object."""
    if six is None:
        lib().orc_set_external_fields(None)
        return
    e = ExtFields()
    e.on = 1
    for c, d in enumerate(six):
        f = e.f[c]
        f.type = int(d.get("type", 0))
        f.value = float(d.get("value", 0.0))
        f.constant = float(d.get("constant", 0.0))
        for k in ("L", "mode", "phase", "C", "A", "X0", "eps"):
            v = d.get(k, (0.0, 0.0))
            getattr(f, k)[0], getattr(f, k)[1] = float(v[0]), float(v[1] if len(v) > 1 else 0.0)
    lib().orc_set_external_fields(C.byref(e))


def add_external_fields(D, x, Ep, Bp):
    lib().orc_add_external_fields.argtypes = [C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p]
    lib().orc_add_external_fields(D, x.shape[1], _ptr(x), _ptr(Ep), _ptr(Bp))


def gather(g, interp, x, xold, E, B):
    n = x.shape[1]
    Ep = np.zeros((3, n))
    Bp = np.zeros((3, n))
    rc = lib().orc_gather(C.byref(g), interp, n, _ptr(x), _ptr(xold), _fabs3(E), _fabs3(B), _ptr(Ep), _ptr(Bp))
    return rc, Ep, Bp


def boris(v, vold, Ep, Bp, fnorm, cnormDt, by_half):
    out = np.empty_like(v)
    lib().orc_boris(v.shape[1], _ptr(out), _ptr(vold), _ptr(Ep), _ptr(Bp), fnorm, cnormDt, int(by_half))
    return out


PUSH_CYL_CYL, PUSH_SPH_SPH, PUSH_CYL_HYB, PUSH_SPH_HYB = 1, 2, 3, 4


def boris_curvilinear(push_type, vold, Ep, Bp, r_old, virt, fnorm, cnormDt, by_half, anticyclic=False):
    """applyForces_CYL_CYL / SPH_SPH / CYL_HYB / SPH_HYB; virt [2, n] is updated in place (types 1, 2)."""
    f = lib().orc_boris_curvilinear
    f.argtypes = [C.c_int, C.c_long] + [C.c_void_p] * 6 + [C.c_double, C.c_double, C.c_int, C.c_int]
    out = np.zeros_like(vold)
    rc = f(push_type, vold.shape[1], _ptr(out), _ptr(vold), _ptr(Ep), _ptr(Bp), _ptr(r_old), _ptr(virt), fnorm, cnormDt,
           int(by_half), int(anticyclic))
    assert rc == 0
    return out


def deposit_current(g, interp, x, xold, v, w, cnormDt, J):
    n = x.shape[1]
    return lib().orc_deposit_current(C.byref(g), interp, n, _ptr(x), _ptr(xold), _ptr(v), _ptr(w), cnormDt, _fabs3(J))


def set_relativistic(relativistic, higuera_cary=False):
    """Switch the push routines to the reference's RELATIVISTIC_PARTICLES build (global, test library)."""
    lib().orc_set_relativistic(int(relativistic), int(higuera_cary))


def implicit_gamma(upold, upbar):
    f = lib().orc_implicit_gamma
    f.restype = C.c_double
    f.argtypes = [C.c_void_p, C.c_void_p]
    a, b = np.ascontiguousarray(upold, dtype=np.float64), np.ascontiguousarray(upbar, dtype=np.float64)
    return f(_ptr(a), _ptr(b))


def advance_positions_implicit_rel(D, x, xold, v, vold, cnormDt):
    f = lib().orc_advance_positions_implicit_rel
    f.argtypes = [C.c_int, C.c_long] + [C.c_void_p] * 4 + [C.c_double]
    f(D, x.shape[1], _ptr(x), _ptr(xold), _ptr(v), _ptr(vold), cnormDt)


def deposit_current_rel(g, interp, x, xold, v, vold, w, cnormDt, J, from_explicit=False):
    f = lib().orc_deposit_current_rel
    f.argtypes = [C.c_void_p, C.c_int, C.c_long] + [C.c_void_p] * 5 + [C.c_double, C.c_int, C.c_void_p]
    return f(C.byref(g), interp, x.shape[1], _ptr(x), _ptr(xold), _ptr(v), _ptr(vold), _ptr(w), cnormDt,
             int(from_explicit), _fabs3(J))


def deposit_rho(g, interp, x, w, stag, rho):
    st = (C.c_int * 2)(*(list(stag) + [0])[:2])
    f = rho.c()
    lib().orc_deposit_rho(C.byref(g), interp, x.shape[1], _ptr(x), _ptr(w), st, C.byref(f))


def scale_fab(f, D, s):
    cf = f.c()
    lib().orc_scale_fab(C.byref(cf), D, s)


def fold_periodic(f, D, stag, valid_lo, valid_hi, periodic):
    cf = f.c()
    i2 = lambda v: (C.c_int * 2)(*(list(v) + [0])[:2])
    lib().orc_fold_periodic(C.byref(cf), D, i2(stag), i2(valid_lo), i2(valid_hi), i2(periodic))


def binomial_filter(f, D, box_lo, box_hi, stag):
    """SpaceUtils::applyBinomialFilter(FArrayBox&, const Box&) (SpaceUtils.cpp:54-112) on a Fab, numpy restatement:
    Q2 = neighbour sum over the grid box (box_lo .. box_hi + stag), then Q = (2^D Q + Q2) / 4^D in the reference's
    operation order.  The reference's update also touches the ghosts with an uninitialised Q2; here they are left
    as they are."""
    a = f.a
    sl = [slice(box_lo[d] - f.lo[d], box_hi[d] + stag[d] - f.lo[d] + 1) for d in range(D)]
    sh = lambda d, k: slice(sl[d].start + k, sl[d].stop + k)
    if D == 1:
        q2 = a[sh(0, 1)] + a[sh(0, -1)]
        a[sl[0]] = (a[sl[0]] * 2.0 + q2) / 4.0
        return
    q2 = 2.0 * (a[sh(0, 1), sl[1]] + a[sh(0, -1), sl[1]]) + 2.0 * (a[sl[0], sh(1, 1)] + a[sl[0], sh(1, -1)])
    q2 = q2 + a[sh(0, 1), sh(1, 1)]
    q2 = q2 + a[sh(0, 1), sh(1, -1)]
    q2 = q2 + a[sh(0, -1), sh(1, 1)]
    q2 = q2 + a[sh(0, -1), sh(1, -1)]
    a[sl[0], sl[1]] = (a[sl[0], sl[1]] * 4.0 + q2) / 16.0


def advance_particles(g, interpE, x, xold, v, vold, E, B, fnorm, cnormDt, order_swap):
    n = x.shape[1]
    return lib().orc_advance_particles(C.byref(g), interpE, n, _ptr(x), _ptr(xold), _ptr(v), _ptr(vold),
                                       _fabs3(E), _fabs3(B), fnorm, cnormDt, int(order_swap))


def advance_particles_iteratively(g, interpE, x, xold, v, vold, E, B, fnorm, cnormDt, rtol, iter_max):
    n = x.shape[1]
    apply_its = C.c_long(0)
    unconv = C.c_long(0)
    its = np.zeros(n, dtype=np.int32)
    rc = lib().orc_advance_particles_iteratively(
        C.byref(g), interpE, n, _ptr(x), _ptr(xold), _ptr(v), _ptr(vold), _fabs3(E), _fabs3(B),
        fnorm, cnormDt, rtol, iter_max, C.byref(apply_its), C.byref(unconv), _ptr(its))
    return rc, apply_its.value, unconv.value, its


def advance_suborbit_and_set_J(g, interpE, interpJ, x, xold, v, vold, w, nsub, E, B, fnorm, cnormDt, rtol, iter_max, J,
                               from_emjacobian=False, max_suborbits=512):
    """advanceSubOrbitParticlesAndSetJ (bulk container): x, xold, v, vold, nsub are updated in place; J accumulates."""
    f = lib().orc_advance_suborbit_particles_and_set_J
    f.argtypes = ([C.c_void_p, C.c_int, C.c_int, C.c_long] + [C.c_void_p] * 8 + [C.c_double] * 3 + [C.c_int] * 3
                  + [C.c_void_p])
    nsub_c = np.ascontiguousarray(nsub, dtype=np.int32)
    rc = f(C.byref(g), interpE, interpJ, x.shape[1], _ptr(x), _ptr(xold), _ptr(v), _ptr(vold), _ptr(w), _ptr(nsub_c),
           _fabs3(E), _fabs3(B), fnorm, cnormDt, rtol, iter_max, int(from_emjacobian), max_suborbits, _fabs3(J))
    nsub[...] = nsub_c
    return rc


def advance_inflow_and_set_J(g, interpE, interpJ, x, xold, v, vold, w, nsub, E, B, fnorm, cnormDt, rtol, iter_max, J,
                             bdry_dir, bdry_side, from_emjacobian=False, max_suborbits=512):
    """advanceInflowParticlesAndSetJ for the inflow list of boundary (bdry_dir, bdry_side); in place like the above."""
    f = lib().orc_advance_inflow_particles_and_set_J
    f.argtypes = ([C.c_void_p, C.c_int, C.c_int, C.c_long] + [C.c_void_p] * 8 + [C.c_double] * 3 + [C.c_int] * 3
                  + [C.c_void_p, C.c_int, C.c_int])
    nsub_c = np.ascontiguousarray(nsub, dtype=np.int32)
    rc = f(C.byref(g), interpE, interpJ, x.shape[1], _ptr(x), _ptr(xold), _ptr(v), _ptr(vold), _ptr(w), _ptr(nsub_c),
           _fabs3(E), _fabs3(B), fnorm, cnormDt, rtol, iter_max, int(from_emjacobian), max_suborbits, _fabs3(J),
           bdry_dir, bdry_side)
    nsub[...] = nsub_c
    return rc


def fast_particles(g, x, xold):
    flag = np.zeros(x.shape[1], dtype=np.int32)
    f = lib().orc_fast_particles
    f.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p]
    f(C.byref(g), x.shape[1], _ptr(x), _ptr(xold), _ptr(flag))
    return flag


def bin_cells(g, x):
    n = x.shape[1]
    cell = np.zeros((g.D, n), dtype=np.int32)
    lib().orc_bin(C.byref(g), n, _ptr(x), _ptr(cell))
    return cell


def cell_moments(g, x, v, w, mass, volume_scale, lo, hi):
    n = x.shape[1]
    shape = [h - l + 1 for l, h in zip(lo, hi)]
    ncell = int(np.prod(shape))
    dens = np.zeros(ncell)
    mom = np.zeros((3, ncell))
    ene = np.zeros((3, ncell))
    i2 = lambda v_: (C.c_int * 2)(*(list(v_) + [0])[:2])
    lib().orc_cell_moments(C.byref(g), n, _ptr(x), _ptr(v), _ptr(w), mass, volume_scale, i2(lo), i2(hi),
                           _ptr(dens), _ptr(mom), _ptr(ene))
    return dens, mom, ene


def debye_length(species_moments):
    """species_moments: list of (dens, mom, ene, mass, charge)."""
    ncell = species_moments[0][0].size
    acc = np.zeros(ncell)
    for dens, mom, ene, mass, charge in species_moments:
        lib().orc_debye_accumulate(ncell, _ptr(dens), _ptr(mom), _ptr(ene), mass, charge, _ptr(acc))
    lib().orc_debye_finish(ncell, _ptr(acc))
    return acc


def scatter_delta_u(u, costh, sinth, cosphi, sinphi):
    dU = np.zeros(3)
    lib().orc_scatter_delta_u(u[0], u[1], u[2], costh, sinth, cosphi, sinphi, _ptr(dU))
    return dU


def ta_delta_u(vp1, den1, vp2, den2, b90_fact, Clog, dt_sec, gauss, u_theta, u_phi):
    dU = np.zeros(3)
    a = np.ascontiguousarray(vp1, dtype=np.float64)
    b = np.ascontiguousarray(vp2, dtype=np.float64)
    lib().orc_ta_delta_u(_ptr(a), den1, _ptr(b), den2, b90_fact, Clog, dt_sec, gauss, u_theta, u_phi, _ptr(dU))
    return dU


def ta_b90_fact(q1, q2, m1, m2):
    return lib().orc_ta_b90_fact(q1, q2, m1, m2)


def ta_self(cell_start, v, dens, mass, charge, Clog, dt_sec):
    npairs = C.c_long(0)
    cs = np.ascontiguousarray(cell_start, dtype=np.int64)
    lib().orc_ta_self(cs.size - 1, _ptr(cs), _ptr(v), v.shape[1], _ptr(dens), mass, charge, Clog, dt_sec, C.byref(npairs))
    return npairs.value


def ta_inter(cs1, v1, dens1, m1, q1, cs2, v2, dens2, m2, q2, Clog, dt_sec):
    npairs = C.c_long(0)
    cs1 = np.ascontiguousarray(cs1, dtype=np.int64)
    cs2 = np.ascontiguousarray(cs2, dtype=np.int64)
    lib().orc_ta_inter(cs1.size - 1, _ptr(cs1), _ptr(v1), v1.shape[1], _ptr(dens1), m1, q1,
                       _ptr(cs2), _ptr(v2), v2.shape[1], _ptr(dens2), m2, q2, Clog, dt_sec, C.byref(npairs))
    return npairs.value


# ---- Coulomb (PROBABILISTIC) and Elastic ------------------------------------------------------
_COUL_SET = False


def _coul_sigs():
    global _COUL_SET
    if _COUL_SET:
        return
    L, dbl, vp, i32, lng = lib(), C.c_double, C.c_void_p, C.c_int, C.c_long
    L.orc_nanbu_costh_sinth.argtypes = [dbl, dbl, vp, vp]
    L.orc_coulomb_delta_u.argtypes = [vp, vp, dbl, dbl, dbl, dbl, dbl, dbl, i32, dbl, dbl, dbl, dbl, dbl, dbl, dbl, vp, vp]
    L.orc_coulomb_delta_u.restype = i32
    L.orc_coulomb_intra.argtypes = [lng, vp, vp, vp, lng, vp, vp, dbl, dbl, dbl, dbl, i32, i32, i32, dbl, vp]
    L.orc_coulomb_inter.argtypes = ([lng, vp, vp, vp, lng, vp, dbl, dbl, vp, vp, vp, lng, vp, dbl, dbl, vp, dbl, dbl,
                                     i32, i32, i32, dbl, vp])
    L.orc_elastic_sigma.argtypes = [dbl, dbl, dbl, i32, vp, vp, vp, i32, i32, vp]
    L.orc_elastic_sigma.restype = dbl
    L.orc_elastic.argtypes = [lng, vp, vp, vp, lng, dbl, vp, vp, vp, lng, vp, dbl, dbl, i32, vp, vp, vp, i32, i32, dbl, vp]
    L.orc_gammainc_3half.restype = dbl
    L.orc_gammainc_3half.argtypes = [dbl]
    L.orc_ta_nu_max.restype = dbl
    L.orc_ta_nu_max.argtypes = [lng, vp, vp, vp, vp, dbl, dbl, dbl, dbl, dbl, i32]
    L.orc_coulomb_nu_max.restype = dbl
    L.orc_coulomb_nu_max.argtypes = [lng, vp, vp, vp, vp, vp, vp, vp, dbl, dbl, dbl, dbl, dbl, i32]
    L.orc_elastic_nu_max.restype = dbl
    L.orc_elastic_nu_max.argtypes = [lng, vp, vp, vp, vp, dbl, dbl, dbl, i32, vp, vp, vp, i32, i32]
    _COUL_SET = True


def nanbu_costh_sinth(s12, U):
    _coul_sigs()
    c, s = C.c_double(0), C.c_double(0)
    lib().orc_nanbu_costh_sinth(s12, U, C.byref(c), C.byref(s))
    return c.value, s.value


def coulomb_delta_u(vp1, vp2, q1, q2, m1, m2, EF_norm, Clog, angular, den12, bmax, sigma_max, dt_sec, gauss, upol, uphi):
    _coul_sigs()
    a, b = np.ascontiguousarray(vp1, dtype=np.float64), np.ascontiguousarray(vp2, dtype=np.float64)
    dU, s12 = np.zeros(3), C.c_double(0)
    lib().orc_coulomb_delta_u(_ptr(a), _ptr(b), q1, q2, m1, m2, EF_norm, Clog, angular, den12, bmax, sigma_max, dt_sec,
                              gauss, upol, uphi, _ptr(dU), C.byref(s12))
    return dU, s12.value


def coulomb_lorentz_scatter(up1, up2, scatter2, q1, q2, m1, m2, EF_norm, Clog, angular, den12, bmax, sigma_max, dt_sec,
                            gauss, upol, uphi):
    """Coulomb::LorentzScatter for one pair; returns (up1', up2', live, s12)."""
    f = lib().orc_coulomb_lorentz_scatter
    f.argtypes = ([C.c_void_p, C.c_void_p, C.c_int] + [C.c_double] * 6 + [C.c_int] + [C.c_double] * 7 + [C.c_void_p])
    a = np.array(up1, dtype=np.float64)
    b = np.array(up2, dtype=np.float64)
    s12 = C.c_double(0)
    live = f(_ptr(a), _ptr(b), int(scatter2), q1, q2, m1, m2, EF_norm, Clog, angular, den12, bmax, sigma_max, dt_sec,
             gauss, upol, uphi, C.byref(s12))
    return a, b, live, s12.value


def coulomb_set_enforce(on, energy_fraction=0.05, energy_fraction_max=0.5, beta_weight_exponent=1, sort_weighted=False,
                        nmin_save=100000):
    """scattering.coulomb.enforce_conservations and companions (Coulomb.H:286-293); global like the deck keys."""
    f = lib().orc_coulomb_set_enforce
    f.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int]
    f(int(on), energy_fraction, energy_fraction_max, int(beta_weight_exponent), int(sort_weighted), int(nmin_save))


def coulomb_set_large_angle(on, test_draw=0.5):
    """scattering.coulomb.include_large_angle_scattering; test_draw = the RL of the explicit-draw entry points"""
    f = lib().orc_coulomb_set_large_angle
    f.argtypes = [C.c_int, C.c_double]
    f.restype = None
    f(int(on), float(test_draw))


def coulomb_set_fas_draws(second=0.5, third=0.5):
    """the second and third uniform of NANBU_FAS / NANBU_FAS_v2 in the explicit-draw entry points (the first is u_polar)"""
    f = lib().orc_coulomb_set_fas_draws
    f.argtypes = [C.c_double, C.c_double]
    f.restype = None
    f(float(second), float(third))


def nanbu_fas_costh_sinth(variant, s12, Clog, b0, bmin_qm, sigma_eff, u1, u2, u3):
    """Coulomb::setNANBUFAScosthsinth (variant 3) / setNANBUFAS_v2_costhsinth (variant 4) with explicit draws"""
    f = lib().orc_nanbu_fas_costh_sinth
    f.argtypes = [C.c_int] + [C.c_double] * 8 + [C.c_void_p, C.c_void_p]
    f.restype = None
    c, sn = C.c_double(), C.c_double()
    f(int(variant), s12, Clog, b0, bmin_qm, sigma_eff, u1, u2, u3, C.byref(c), C.byref(sn))
    return c.value, sn.value


def coulomb_set_weight_method(conservative):
    """Coulomb weight_method: False PROBABILISTIC, True CONSERVATIVE (Sentoku-Kemp, Coulomb.cpp:730-917, 1439-1640)."""
    f = lib().orc_coulomb_set_weight_method
    f.argtypes = [C.c_int]
    f(int(conservative))


def coulomb_intra(cell_start, v, w, dens, LDe, cellV_SI, mass, charge, Clog, angular, NxN, NxN_Nthresh, dt_sec):
    _coul_sigs()
    npairs = C.c_long(0)
    cs = np.ascontiguousarray(cell_start, dtype=np.int64)
    lib().orc_coulomb_intra(cs.size - 1, _ptr(cs), _ptr(v), _ptr(w), v.shape[1], _ptr(dens), _ptr(LDe), cellV_SI, mass,
                            charge, Clog, angular, int(NxN), NxN_Nthresh, dt_sec, C.byref(npairs))
    return npairs.value


def coulomb_inter(cs1, v1, w1, dens1, m1, q1, cs2, v2, w2, dens2, m2, q2, LDe, cellV_SI, Clog, angular, NxN,
                  NxN_Nthresh, dt_sec):
    _coul_sigs()
    npairs = C.c_long(0)
    cs1 = np.ascontiguousarray(cs1, dtype=np.int64)
    cs2 = np.ascontiguousarray(cs2, dtype=np.int64)
    lib().orc_coulomb_inter(cs1.size - 1, _ptr(cs1), _ptr(v1), _ptr(w1), v1.shape[1], _ptr(dens1), m1, q1, _ptr(cs2),
                            _ptr(v2), _ptr(w2), v2.shape[1], _ptr(dens2), m2, q2, _ptr(LDe), cellV_SI, Clog, angular,
                            int(NxN), NxN_Nthresh, dt_sec, C.byref(npairs))
    return npairs.value


def elastic_sigma(g12, mu, const_sigma=0.0, E=None, Q=None, XI=None, angular=0, loglog=False):
    _coul_sigs()
    xi = C.c_double(0)
    n = 0 if E is None else len(E)
    arr = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (E, Q, XI)]
    s = lib().orc_elastic_sigma(g12, mu, const_sigma, n, *[None if a is None else _ptr(a) for a in arr], angular,
                                int(loglog), C.byref(xi))
    return s, xi.value


def elastic(cs1, v1, w1, m1, cs2, v2, w2, dens2, m2, dt_sec, const_sigma=0.0, E=None, Q=None, XI=None, angular=0,
            loglog=False):
    _coul_sigs()
    ncoll = C.c_long(0)
    cs1 = np.ascontiguousarray(cs1, dtype=np.int64)
    cs2 = np.ascontiguousarray(cs2, dtype=np.int64)
    n = 0 if E is None else len(E)
    arr = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (E, Q, XI)]
    lib().orc_elastic(cs1.size - 1, _ptr(cs1), _ptr(v1), _ptr(w1), v1.shape[1], m1, _ptr(cs2), _ptr(v2), _ptr(w2),
                      v2.shape[1], _ptr(dens2), m2, const_sigma, n, *[None if a is None else _ptr(a) for a in arr],
                      angular, int(loglog), dt_sec, C.byref(ncoll))
    return ncoll.value


def elastic_conservative(cs1, v1, w1, m1, cs2, v2, w2, dens2, m2, dt_sec, const_sigma):
    """Elastic::electronImpact with weight_method = CONSERVATIVE, constant cross section; v1, v2 and w2 change."""
    f = lib().orc_elastic_wm
    f.argtypes = ([C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                   C.c_long, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                   C.c_int, C.c_int, C.c_double, C.c_void_p])
    ncoll = C.c_long(0)
    a1 = np.ascontiguousarray(cs1, dtype=np.int64)
    a2 = np.ascontiguousarray(cs2, dtype=np.int64)
    f(a1.size - 1, _ptr(a1), _ptr(v1), _ptr(w1), v1.shape[1], m1, _ptr(a2), _ptr(v2), _ptr(w2), v2.shape[1], _ptr(dens2),
      m2, const_sigma, 0, None, None, None, 0, 0, 1, dt_sec, C.byref(ncoll))
    return ncoll.value


def ta_nu_max(m1, m2, q1, q2, mass1, mass2, Clog, intra):
    """TakizukaAbe::setMeanFreeTime: m = (dens[ncell], mom[3, ncell], ene[3, ncell]) per species."""
    _coul_sigs()
    d1, _, e1 = [np.ascontiguousarray(a, dtype=np.float64) for a in m1]
    d2, _, e2 = [np.ascontiguousarray(a, dtype=np.float64) for a in m2]
    return lib().orc_ta_nu_max(d1.size, _ptr(d1), _ptr(e1), _ptr(d2), _ptr(e2), q1, q2, mass1, mass2, Clog, int(intra))


def coulomb_nu_max(LDe, m1, m2, q1, q2, mass1, mass2, Clog, intra):
    _coul_sigs()
    LDe = np.ascontiguousarray(LDe, dtype=np.float64)
    d1, p1, e1 = [np.ascontiguousarray(a, dtype=np.float64) for a in m1]
    d2, p2, e2 = [np.ascontiguousarray(a, dtype=np.float64) for a in m2]
    return lib().orc_coulomb_nu_max(d1.size, _ptr(LDe), _ptr(d1), _ptr(p1), _ptr(e1), _ptr(d2), _ptr(p2), _ptr(e2),
                                    q1, q2, mass1, mass2, Clog, int(intra))


def elastic_nu_max(m1, m2, mass1, mass2, const_sigma=0.0, E=None, Q=None, XI=None, angular=0, loglog=False):
    _coul_sigs()
    d1, _, e1 = [np.ascontiguousarray(a, dtype=np.float64) for a in m1]
    d2, _, e2 = [np.ascontiguousarray(a, dtype=np.float64) for a in m2]
    n = 0 if E is None else len(E)
    arr = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (E, Q, XI)]
    return lib().orc_elastic_nu_max(d1.size, _ptr(d1), _ptr(e1), _ptr(d2), _ptr(e2), mass1, mass2, const_sigma, n,
                                    *[None if a is None else _ptr(a) for a in arr], angular, int(loglog))


def gammainc_3half(x):
    _coul_sigs()
    return lib().orc_gammainc_3half(float(x))


# ---- mass matrices (oracle_massmatrix.cpp) ---------------------------------------------------
SIGMA_NAMES = ["xx", "xy", "xz", "yx", "yy", "yz", "zx", "zy", "zz"]


class MFab:
    """ncomp components of one box in CHF_FRA layout (column major, component index slowest)."""

    def __init__(self, lo, hi, ncomp):
        self.lo = tuple(int(v) for v in lo)
        self.hi = tuple(int(v) for v in hi)
        self.ncomp = int(ncomp)
        shape = tuple(h - l + 1 for l, h in zip(self.lo, self.hi)) + (self.ncomp,)
        self.a = np.zeros(shape, dtype=np.float64, order="F")

    def c(self):
        f = CMFab()
        f.p = self.a.ctypes.data_as(C.POINTER(C.c_double))
        for d in range(2):
            f.lo[d] = self.lo[d] if d < len(self.lo) else 0
            f.hi[d] = self.hi[d] if d < len(self.hi) else 0
        f.ncomp = self.ncomp
        return f


def mm_ncomp(D, interp, ghosts):
    """(9, 2) int array: components per direction of sigma_xx .. sigma_zz."""
    nc = np.zeros((9, 2), dtype=np.int32)
    rc = lib().orc_mm_ncomp(D, interp, ghosts, _ptr(nc))
    if rc:
        raise ValueError("the reference asserts against this interp/ghosts/D combination")
    return nc


def j_stag(D):
    """centring of Jx, Jy, Jz (= that of Ex, Ey, Ez)"""
    return E_STAG[D]


def mm_alloc(D, interp, ghosts, box_lo, box_hi):
    """Nine zeroed sigma containers on the boxes of the J component of their row."""
    nc = mm_ncomp(D, interp, ghosts)
    out = []
    for k in range(9):
        stag = E_STAG[D][k // 3]
        lo = [l - ghosts for l in box_lo]
        hi = [h + ghosts + s for h, s in zip(box_hi, stag)]
        out.append(MFab(lo, hi, int(nc[k, 0]) * int(nc[k, 1])))
    return nc, out


def _mfabs9(sig):
    arr = (CMFab * 9)()
    for i, f in enumerate(sig):
        arr[i] = f.c()
    return arr


def mm_kernels(Bp, qp, alphas, volume, upold, upbar, anticyclic=1, relativistic=False):
    out = np.zeros(12)
    Bp = np.ascontiguousarray(Bp, dtype=np.float64)
    upold = np.ascontiguousarray(upold, dtype=np.float64)
    upbar = np.ascontiguousarray(upbar, dtype=np.float64)
    f = lib().orc_mm_kernels
    f.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    f(_ptr(Bp), qp, alphas, volume, _ptr(upold), _ptr(upbar), int(anticyclic), int(relativistic), _ptr(out))
    return out[:3].copy(), out[3:].reshape(3, 3).copy()


def deposit_mass_matrices(g, interp, x, xold, v, vold, w, qovs, alphas, cnormDt, B, J0, sigma, anticyclic=False,
                          relativistic=False):
    f = lib().orc_deposit_mass_matrices
    f.argtypes = ([C.c_void_p, C.c_int, C.c_long] + [C.c_void_p] * 5 + [C.c_double] * 3 + [C.c_int, C.c_int] +
                  [C.c_void_p] * 3)
    n = x.shape[1]
    return f(C.byref(g), interp, n, _ptr(x), _ptr(xold), _ptr(v), _ptr(vold), _ptr(w), qovs, alphas, cnormDt,
             int(anticyclic), int(relativistic), _fabs3(B), _fabs3(J0), _mfabs9(sigma))


def compute_J_from_mass_matrices(D, ncomp, sigma, E0, E, J0, J):
    f = lib().orc_compute_J_from_mass_matrices
    f.argtypes = [C.c_int] + [C.c_void_p] * 6
    nc = np.ascontiguousarray(ncomp, dtype=np.int32)
    f(D, _ptr(nc), _mfabs9(sigma), _fabs3(E0), _fabs3(E), _fabs3(J0), _fabs3(J))


# ---- HardSphere (no-time-counter) -------------------------------------------------------------
def hs_sigmaT(r1, r2):
    f = lib().orc_hs_sigmaT
    f.restype = C.c_double
    f.argtypes = [C.c_double, C.c_double]
    return f(r1, r2)


def hs_self(cell_start, v, w, dens, ene, mass, sigmaT, dt_sec):
    """HardSphere::applySelfScattering; v [3][n] is updated in place.  Returns (candidates, collisions)."""
    f = lib().orc_hs_self
    f.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p] + [C.c_double] * 3 + \
                 [C.c_void_p] * 2
    cs = np.ascontiguousarray(cell_start, dtype=np.int64)
    ene = np.ascontiguousarray(ene, dtype=np.float64)
    a, b = C.c_long(0), C.c_long(0)
    f(cs.size - 1, _ptr(cs), _ptr(v), _ptr(w), v.shape[1], _ptr(dens), _ptr(ene), mass, sigmaT, dt_sec, C.byref(a),
      C.byref(b))
    return a.value, b.value


def hs_self_conservative(cell_start, v, w, dens, ene, mass, sigmaT, dt_sec):
    """HardSphere::applySelfScattering with weight_method = CONSERVATIVE; v and w are updated in place."""
    f = lib().orc_hs_self_wm
    f.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                  C.c_int, C.c_double, C.c_void_p, C.c_void_p]
    cs = np.ascontiguousarray(cell_start, dtype=np.int64)
    ene = np.ascontiguousarray(ene, dtype=np.float64)
    a, b = C.c_long(0), C.c_long(0)
    f(cs.size - 1, _ptr(cs), _ptr(v), _ptr(w), v.shape[1], _ptr(dens), _ptr(ene), mass, sigmaT, 1, dt_sec, C.byref(a),
      C.byref(b))
    return a.value, b.value


def hs_inter(cs1, v1, w1, dens1, ene1, m1, cs2, v2, w2, dens2, ene2, m2, Vc, sigmaT, dt_sec):
    f = lib().orc_hs_inter
    f.argtypes = ([C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_double,
                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p] + [C.c_double] * 4 +
                  [C.c_void_p] * 2)
    a1 = np.ascontiguousarray(cs1, dtype=np.int64)
    a2 = np.ascontiguousarray(cs2, dtype=np.int64)
    e1, e2 = np.ascontiguousarray(ene1, dtype=np.float64), np.ascontiguousarray(ene2, dtype=np.float64)
    a, b = C.c_long(0), C.c_long(0)
    f(a1.size - 1, _ptr(a1), _ptr(v1), _ptr(w1), v1.shape[1], _ptr(dens1), _ptr(e1), m1, _ptr(a2), _ptr(v2), _ptr(w2),
      v2.shape[1], _ptr(dens2), _ptr(e2), m2, Vc, sigmaT, dt_sec, C.byref(a), C.byref(b))
    return a.value, b.value


def hs_inter_conservative(cs1, v1, w1, dens1, ene1, m1, cs2, v2, w2, dens2, ene2, m2, Vc, sigmaT, dt_sec):
    """HardSphere between species, weight_method CONSERVATIVE: v1, w1, v2, w2 change in place"""
    f = lib().orc_hs_inter_wm
    f.argtypes = ([C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_double,
                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p] + [C.c_double] * 3 + [C.c_int] +
                  [C.c_double] + [C.c_void_p] * 2)
    a1 = np.ascontiguousarray(cs1, dtype=np.int64)
    a2 = np.ascontiguousarray(cs2, dtype=np.int64)
    e1, e2 = np.ascontiguousarray(ene1, dtype=np.float64), np.ascontiguousarray(ene2, dtype=np.float64)
    a, b = C.c_long(0), C.c_long(0)
    f(a1.size - 1, _ptr(a1), _ptr(v1), _ptr(w1), v1.shape[1], _ptr(dens1), _ptr(e1), m1, _ptr(a2), _ptr(v2), _ptr(w2),
      v2.shape[1], _ptr(dens2), _ptr(e2), m2, Vc, sigmaT, 1, dt_sec, C.byref(a), C.byref(b))
    return a.value, b.value


def vhs_consts(mass, eta, T0, mu0):
    """(4 pi A, 4/alpha) of VariableHardSphere::initialize"""
    a, b = C.c_double(0), C.c_double(0)
    f = lib().orc_vhs_consts
    f.argtypes = [C.c_double] * 4 + [C.c_void_p] * 2
    f(mass, eta, T0, mu0, C.byref(a), C.byref(b))
    return a.value, b.value


def vhs_self(cell_start, v, dens, ene, mass, fourPiA, fourOverAlpha, dt_sec):
    f = lib().orc_vhs_self
    f.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p] + [C.c_double] * 4 + [C.c_void_p] * 2
    cs = np.ascontiguousarray(cell_start, dtype=np.int64)
    ene = np.ascontiguousarray(ene, dtype=np.float64)
    a, b = C.c_long(0), C.c_long(0)
    f(cs.size - 1, _ptr(cs), _ptr(v), v.shape[1], _ptr(dens), _ptr(ene), mass, fourPiA, fourOverAlpha, dt_sec, C.byref(a),
      C.byref(b))
    return a.value, b.value


# ---- AoS + linked-list driver of the same arithmetic (oracle_aos.cpp) ---------------------------
class AosList:
    """Particles as 176-byte objects in a doubly linked list, advanced and deposited with one kernel call per
    particle -- the reference's memory behaviour (bench.py's faithful CPU number)."""

    def __init__(self, D, x, xold, v, vold, w, scattered=False):
        L = lib()
        L.orc_aos_create_ex.restype = C.c_void_p
        L.orc_aos_create_ex.argtypes = [C.c_int, C.c_long] + [C.c_void_p] * 5 + [C.c_int]
        L.orc_aos_destroy.argtypes = [C.c_void_p]
        L.orc_aos_advance_deposit.argtypes = ([C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p] + [C.c_double] * 3 +
                                              [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p])
        L.orc_aos_read.argtypes = [C.c_void_p] * 3
        self.D, self.n = D, x.shape[1]
        self.h = L.orc_aos_create_ex(D, self.n, _ptr(x), _ptr(xold), _ptr(v), _ptr(vold), _ptr(w), int(scattered))
        if not self.h:
            raise RuntimeError("orc_aos_create failed (relativistic build is not supported)")

    def advance_deposit(self, g, interp, E, B, fnorm, cnormDt, rtol, iter_max, J):
        a, b = C.c_long(0), C.c_long(0)
        rc = lib().orc_aos_advance_deposit(self.h, C.byref(g), interp, _fabs3(E), _fabs3(B), fnorm, cnormDt, rtol, iter_max,
                                           _fabs3(J), C.byref(a), C.byref(b))
        return rc, a.value, b.value

    def read(self):
        x, v = np.zeros((self.D, self.n)), np.zeros((3, self.n))
        lib().orc_aos_read(self.h, _ptr(x), _ptr(v))
        return x, v

    def destroy(self):
        if self.h:
            lib().orc_aos_destroy(self.h)
            self.h = None
