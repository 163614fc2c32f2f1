#!/usr/bin/env python
"""Generates tests/golden/ref_pins_collapse.npz: ScatteringUtils::collapseThreeToTwo of the REFERENCE
(oracle/_ref/libpicnic_ref.so, built by oracle/ref_build.sh from /root/reference) on seeded inputs."""
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
    lib.ref_collapse_three_to_two.argtypes = [C.c_void_p] * 5 + [C.c_double]
    return lib


def inputs(seed=20261018, n=200):
    rng = np.random.default_rng(seed)
    d = {"vp2": rng.standard_normal((n, 3)) * 0.01, "vp3": rng.standard_normal((n, 3)) * 0.01,
         "wp3": rng.random(n) + 0.5}
    d["wp2"] = rng.random(n) + 1.0
    d["wp2p"] = d["wp2"] * rng.uniform(0.05, 0.95, n)            # CH_assert(wp2 > wp2p)
    d["vp2p"] = d["vp2"] + rng.standard_normal((n, 3)) * 0.005   # the scattered fraction of particle 2
    return d


def run(fn, d):
    n = d["wp2"].size
    o2, o3, w2, w3 = d["vp2"].copy(), d["vp3"].copy(), d["wp2"].copy(), d["wp3"].copy()
    for i in range(n):
        a, b = np.ascontiguousarray(o2[i]), np.ascontiguousarray(o3[i])
        wa, wb = C.c_double(w2[i]), C.c_double(w3[i])
        c = np.ascontiguousarray(d["vp2p"][i])
        fn(a.ctypes.data, C.byref(wa), b.ctypes.data, C.byref(wb), c.ctypes.data, float(d["wp2p"][i]))
        o2[i], o3[i], w2[i], w3[i] = a, b, wa.value, wb.value
    return o2, o3, w2, w3


if __name__ == "__main__":
    d = inputs()
    o2, o3, w2, w3 = run(ref_lib().ref_collapse_three_to_two, d)
    out = {("in_" + k): v for k, v in d.items()}
    out.update(out_vp2=o2, out_vp3=o3, out_wp2=w2, out_wp3=w3)
    np.savez(os.path.join(HERE, "ref_pins_collapse.npz"), **out)
    print("wrote ref_pins_collapse.npz", o2.shape)
