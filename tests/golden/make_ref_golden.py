#!/usr/bin/env python
"""Generates tests/golden/ref_pins.npz by running the REFERENCE's own code
(oracle/_ref/libpicnic_ref.so, built by oracle/ref_build.sh from /root/reference against the Chombo
mock) on seeded inputs.  Run in the container that has /root/reference; the .npz is committed so
that the pins travel to boxes without the reference.

Pinned: PicSpeciesUtils::applyForces (Boris, both byHalfDt), ScatteringUtils::computeDeltaU,
rotateVelocity, getScatteringCos, JustinsParticle::linearOut (wire format)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def ref_lib():
    so = os.path.join(ROOT, "oracle", "_ref", "libpicnic_ref.so")
    if not os.path.exists(so):
        subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "ref_build.sh")])
    lib = C.CDLL(so)
    dbl = C.c_double
    lib.ref_boris.argtypes = [C.c_long] + [C.c_void_p] * 4 + [dbl, dbl, C.c_int]
    lib.ref_delta_u.argtypes = [dbl] * 7 + [C.c_void_p]
    lib.ref_rotate_velocity.argtypes = [C.c_void_p] + [dbl] * 4
    lib.ref_scattering_cos.argtypes = [dbl, dbl]
    lib.ref_scattering_cos.restype = dbl
    lib.ref_particle_wire.argtypes = [dbl] + [C.c_void_p] * 4 + [C.c_ulonglong, C.c_void_p]
    return lib


def inputs(seed=20261017, n=96):
    rng = np.random.default_rng(seed)
    d = {}
    d["vold"] = rng.standard_normal((3, n)) * 0.05
    d["Ep"] = rng.standard_normal((3, n)) * 3.0
    d["Bp"] = rng.standard_normal((3, n)) * 2.0
    d["Bp"][:, :4] = 0.0                       # unmagnetised corner
    d["fnorm"], d["cnormDt"] = -0.731, 0.213
    u = rng.standard_normal((3, n)) * 0.02
    u[0, :6] = 0.0
    u[1, :6] = 0.0                             # uperp == 0 branch
    u[2, 3:6] *= -1.0
    d["u"] = u
    th, ph = rng.random(n) * np.pi, rng.random(n) * 2 * np.pi
    d["costh"], d["sinth"], d["cosphi"], d["sinphi"] = np.cos(th), np.sin(th), np.cos(ph), np.sin(ph)
    d["R"], d["xi"] = rng.random(n), rng.random(n)
    return d


def run_reference(d):
    lib = ref_lib()
    n = d["vold"].shape[1]
    out = {}
    p = lambda a: a.ctypes.data
    for half in (0, 1):
        v = np.zeros((3, n))
        lib.ref_boris(n, p(v), p(np.ascontiguousarray(d["vold"])), p(np.ascontiguousarray(d["Ep"])),
                      p(np.ascontiguousarray(d["Bp"])), d["fnorm"], d["cnormDt"], half)
        out["boris_half%d" % half] = v
    dU = np.zeros((n, 3))
    rot = np.zeros((n, 3))
    cs = np.zeros(n)
    for i in range(n):
        t = np.zeros(3)
        lib.ref_delta_u(d["u"][0, i], d["u"][1, i], d["u"][2, i], d["costh"][i], d["sinth"][i], d["cosphi"][i],
                        d["sinphi"][i], p(t))
        dU[i] = t
        t = np.ascontiguousarray(d["u"][:, i].copy())
        lib.ref_rotate_velocity(p(t), d["costh"][i], d["sinth"][i], d["cosphi"][i], d["sinphi"][i])
        rot[i] = t
        cs[i] = lib.ref_scattering_cos(d["R"][i], d["xi"][i])
    out["delta_u"], out["rotate"], out["scatter_cos"] = dU, rot, cs
    # wire format of one 2D particle
    buf = np.zeros(32)
    x, xo = np.array([1.25, -3.5]), np.array([1.0, -3.25])
    v, vo = np.array([0.1, 0.2, 0.3]), np.array([0.4, 0.5, 0.6])
    nbytes = lib.ref_particle_wire(7.5, p(x), p(xo), p(v), p(vo), 123456789, p(buf))
    out["wire"] = buf[:nbytes // 8].copy()
    return out


def run_reference_relativistic(seed=20261018, n=96):
    """The reference compiled with -DRELATIVISTIC_PARTICLES (oracle/_ref/libpicnic_ref_rel.so): applyForces with the
    Boris and the Higuera-Cary gamma, both byHalfDt, and PicSpeciesUtils::getImplicitGamma, on velocities up to
    gamma*beta ~ 3."""
    so = os.path.join(ROOT, "oracle", "_ref", "libpicnic_ref_rel.so")
    if not os.path.exists(so):
        subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "ref_build.sh")])
    lib = C.CDLL(so)
    dbl = C.c_double
    lib.ref_boris.argtypes = [C.c_long] + [C.c_void_p] * 4 + [dbl, dbl, C.c_int]
    lib.ref_implicit_gamma.argtypes = [C.c_void_p, C.c_void_p]
    lib.ref_implicit_gamma.restype = dbl
    assert lib.ref_is_relativistic() == 1
    rng = np.random.default_rng(seed)
    d = {"vold": rng.standard_normal((3, n)) * 1.2, "Ep": rng.standard_normal((3, n)) * 3.0,
         "Bp": rng.standard_normal((3, n)) * 2.0, "fnorm": -0.731, "cnormDt": 0.213,
         "ubar": rng.standard_normal((3, n)) * 1.2}
    d["Bp"][:, :4] = 0.0
    d["vold"][:, 4:8] *= 1e-3                       # non-relativistic corner
    p = lambda a: a.ctypes.data
    out = {}
    for hc in (0, 1):
        lib.ref_set_higuera_cary(hc)
        for half in (0, 1):
            v = np.zeros((3, n))
            lib.ref_boris(n, p(v), p(np.ascontiguousarray(d["vold"])), p(np.ascontiguousarray(d["Ep"])),
                          p(np.ascontiguousarray(d["Bp"])), d["fnorm"], d["cnormDt"], half)
            out["boris_hc%d_half%d" % (hc, half)] = v
    g = np.zeros(n)
    for i in range(n):
        uo, ub = np.ascontiguousarray(d["vold"][:, i]), np.ascontiguousarray(d["ubar"][:, i])   # keep them alive
        g[i] = lib.ref_implicit_gamma(p(uo), p(ub))
    out["implicit_gamma"] = g
    return d, out


if __name__ == "__main__":
    dr, outr = run_reference_relativistic()
    np.savez(os.path.join(HERE, "ref_pins_rel.npz"), **{"in_" + k: np.asarray(v) for k, v in dr.items()},
             **{"out_" + k: v for k, v in outr.items()})
    print("wrote ref_pins_rel.npz:", {k: v.shape for k, v in outr.items()})
    d = inputs()
    out = run_reference(d)
    np.savez(os.path.join(HERE, "ref_pins.npz"), **{"in_" + k: np.asarray(v) for k, v in d.items()},
             **{"out_" + k: v for k, v in out.items()})
    print("wrote ref_pins.npz:", {k: v.shape for k, v in out.items()})
