#!/usr/bin/env python
"""Generates tests/golden/ref_pins_gather.npz by running the REFERENCE's own C++ gathers
(MeshInterp::interpolateEMfieldsToPart_testing -> interpolateEMfieldsToPart_CIC / _TSC, interpolateBfieldsToPart_CIC,
interpolateEToPart_CC0 1D + 2D, interpolateEToPart_CC1 1D; src/particle_tools/MeshInterpI.H:1013-1850) from
oracle/_ref/libpicnic_ref_mi{1,2}d.so, which oracle/ref_build.sh compiles from /root/reference against the Chombo
mock.  Run in the container that has /root/reference; the .npz is committed so the pins travel.

Cases: D in {1, 2} x {CIC, TSC, CC0} and CC1 in 1D (the reference's 2D C++ CC1 routine is, in its own words, "just a
copy of _CC0 in 2D" and is not dispatched).  Inputs are seeded; orbits span one to three segments."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_INTERP = {"CIC": 2, "TSC": 3, "CC0": 5, "CC1": 6}     # MeshInterp.H:27 InterpType
E_STAG = {1: [(0,), (1,), (1,)], 2: [(0, 1), (1, 0), (1, 1)]}
B_STAG = {1: [(1,), (0,), (0,)], 2: [(1, 0), (0, 1), (0, 0)]}
CASES = [(1, "CIC"), (1, "TSC"), (1, "CC0"), (1, "CC1"), (2, "CIC"), (2, "TSC"), (2, "CC0")]
NCELL, GHOSTS, DX, XMIN = 12, 3, (0.25, 0.2), (-0.5, 0.3)


def ref_lib(D):
    so = os.path.join(ROOT, "oracle", "_ref", "libpicnic_ref_mi%dd.so" % D)
    if not os.path.exists(so):
        subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "ref_build.sh")])
    lib = C.CDLL(so)
    lib.refmi_gather.argtypes = [C.c_int, C.c_long] + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 7
    assert lib.refmi_spacedim() == D
    return lib


def inputs(D, seed):
    """Fields on the ghosted arrays of a box of NCELL cells per direction; particles well inside the box so that
    every stencil of a three-segment orbit stays inside the arrays."""
    rng = np.random.default_rng(seed)
    lo_box, hi_box = [0] * D, [NCELL - 1] * D
    fields = []
    for stag in E_STAG[D] + B_STAG[D]:
        lo = [l - GHOSTS for l in lo_box]
        hi = [h + GHOSTS + s for h, s in zip(hi_box, stag)]
        shape = tuple(h - l + 1 for l, h in zip(lo, hi))
        fields.append((lo, hi, stag, np.asfortranarray(rng.standard_normal(shape) * 3.0 + 1.0)))
    n = 400
    dx, le = np.array(DX[:D]), np.array(XMIN[:D])
    xold = le[:, None] + dx[:, None] * rng.uniform(2.0, NCELL - 2.0, size=(D, n))
    # half displacements of 0 .. 0.7 cells: x_new = xold + 2 dxp is up to 1.4 cells away (1-3 segments per direction)
    scale = np.where(rng.random(n) < 0.5, 0.2, 0.7)
    dxp = dx[:, None] * rng.uniform(-1.0, 1.0, size=(D, n)) * scale
    dxp[:, :8] = 0.0                                        # particles at rest (dXp == 0 guards)
    if D == 2:
        dxp[0, 8:16] = 0.0                                  # motion along one axis only (slope 0 / inf)
        dxp[1, 16:24] = 0.0
    # a few particles exactly on faces / nodes / cell centres
    xold[:, 24:28] = le[:, None] + dx[:, None] * np.array([3.0, 4.0, 5.5, 6.5])[None, :]
    x = xold + dxp
    return fields, np.ascontiguousarray(x), np.ascontiguousarray(xold)


def run_reference(D, interp, fields, x, xold):
    lib = ref_lib(D)
    n = x.shape[1]
    le = np.array(XMIN[:D])
    dx = np.array(DX[:D])
    re = le + NCELL * dx
    # 1D: the virtual components arrive as ONE two-component array for (Ey, Ez) and one for (By, Bz)
    # (PicChargedSpecies.cpp:3899-3913); component index slowest
    arrs, lo, hi, typ, ncomp = [], [], [], [], []
    if D == 1:
        ev = np.ascontiguousarray(np.stack([fields[1][3], fields[2][3]]))
        bv = np.ascontiguousarray(np.stack([fields[4][3], fields[5][3]]))
        pick = [(fields[0][3], 0, 1), (ev, 1, 2), (ev, 2, 2), (fields[3][3], 3, 1), (bv, 4, 2), (bv, 5, 2)]
    else:
        pick = [(fields[c][3], c, 1) for c in range(6)]
    for a, c, nc in pick:
        arrs.append(a)
        lo += list(fields[c][0])
        hi += list(fields[c][1])
        typ += list(fields[c][2])
        ncomp.append(nc)
    F = (C.c_void_p * 6)(*[a.ctypes.data for a in arrs])
    lo, hi, typ, ncomp = (np.array(v, dtype=np.int32) for v in (lo, hi, typ, ncomp))
    Ep, Bp = np.zeros((3, n)), np.zeros((3, n))
    p = lambda a: a.ctypes.data
    rc = lib.refmi_gather(REF_INTERP[interp], n, p(x), p(xold), p(le), p(re), p(dx), GHOSTS, F, p(lo), p(hi), p(typ),
                          p(ncomp), p(Ep), p(Bp))
    assert rc == 0
    return Ep, Bp


def main():
    out = {}
    for (D, interp) in CASES:
        fields, x, xold = inputs(D, 1000 * D + REF_INTERP[interp])
        Ep, Bp = run_reference(D, interp, fields, x, xold)
        tag = "%dd_%s" % (D, interp)
        out["x_" + tag], out["xold_" + tag], out["Ep_" + tag], out["Bp_" + tag] = x, xold, Ep, Bp
        for c in range(6):
            out["F%d_%s" % (c, tag)] = fields[c][3]
    np.savez_compressed(os.path.join(HERE, "ref_pins_gather.npz"), **out)
    print("wrote ref_pins_gather.npz:", sorted(k for k in out if k.startswith("Ep_")))


if __name__ == "__main__":
    main()
