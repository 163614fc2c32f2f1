#!/usr/bin/env python
"""Generates tests/golden/ref_pins_curv.npz by running the REFERENCE's own curvilinear velocity pushes
(PicSpeciesUtils::applyForces_CYL_CYL / _SPH_SPH / _CYL_HYB / _SPH_HYB, src/species/pic/PicSpeciesUtils.cpp:103-473, compiled
from /root/reference into oracle/_ref/libpicnic_ref.so and, with -DRELATIVISTIC_PARTICLES, libpicnic_ref_rel.so by
oracle/ref_build.sh) on seeded inputs.  Run in the container that has /root/reference; the .npz is committed."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def ref_lib(rel):
    so = os.path.join(ROOT, "oracle", "_ref", "libpicnic_ref_rel.so" if rel else "libpicnic_ref.so")
    if not os.path.exists(so):
        subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "ref_build.sh")])
    lib = C.CDLL(so)
    lib.ref_boris_curvilinear.argtypes = ([C.c_int, C.c_long] + [C.c_void_p] * 6 + [C.c_double, C.c_double, C.c_int, C.c_int])
    return lib


def inputs(seed=20261018, n=64):
    rng = np.random.default_rng(seed)
    d = {}
    d["vold"] = rng.standard_normal((3, n)) * 0.05
    d["Ep"] = rng.standard_normal((3, n)) * 3.0
    d["Bp"] = rng.standard_normal((3, n)) * 2.0
    d["r_old"] = 0.05 + rng.random(n) * 2.0
    virt = rng.standard_normal((2, n)) * 0.05
    virt[:, : n // 2] = 0.0                     # dtheta == 0: the predictor-corrector branch of CYL_CYL / SPH_SPH
    d["virt"] = virt
    d["fnorm"], d["cnormDt"] = -0.731, 0.213
    return d


def cases():
    for rel in (0, 1):
        for ptype in (1, 2, 3, 4):
            for half in ((0, 1) if ptype in (1, 2) else (1,)):
                for anti in ((0, 1) if ptype in (1, 3) else (0,)):
                    yield rel, ptype, half, anti


def run_reference(d):
    n = d["vold"].shape[1]
    out = {}
    p = lambda a: a.ctypes.data
    for rel, ptype, half, anti in cases():
        lib = ref_lib(rel)
        v = np.zeros((3, n))
        virt = np.ascontiguousarray(d["virt"].copy())
        lib.ref_boris_curvilinear(ptype, n, p(v), p(np.ascontiguousarray(d["vold"])), p(np.ascontiguousarray(d["Ep"])),
                                  p(np.ascontiguousarray(d["Bp"])), p(np.ascontiguousarray(d["r_old"])), p(virt),
                                  d["fnorm"], d["cnormDt"], half, anti)
        key = "r%d_t%d_h%d_a%d" % (rel, ptype, half, anti)
        out["v_" + key] = v
        out["virt_" + key] = virt
    return out


if __name__ == "__main__":
    d = inputs()
    out = run_reference(d)
    np.savez(os.path.join(HERE, "ref_pins_curv.npz"), **{"in_" + k: np.asarray(v) for k, v in d.items()},
             **{"out_" + k: v for k, v in out.items()})
    print("wrote ref_pins_curv.npz with", len(out), "arrays")
