#!/usr/bin/env python
"""Write the inputs of the REFERENCE-RUN golden vectors: tests/golden/ref_inputs.f64 (+ .txt manifest).

The reference (FLOWVPM.jl + FastMultipole.jl) cannot run in this repository's containers (no
Julia), so parity is pinned to mpmath / analytic answers only ("Oracle").  To pin it
to the reference itself, anyone with Julia runs

    julia --project=/path/to/FLOWVPM.jl baseline/julia/make_golden.jl

which reads the inputs written here, runs the reference's own UJ_direct on them and writes
tests/golden/ref_outputs.f64 (+ .txt).  tests/test_reference_golden.py then checks the oracle
(CPU) and the CUDA path (GPU) against those outputs; without the file those tests are skipped.

Format (both files): raw little-endian float64; the manifest has one line per array,
`name rows cols offset_in_doubles`, column-major (Julia's and numpy-'F' order).
Cases: a 40-particle cloud with static particles (the mpmath golden inputs), the C1 ring
(4900 particles), a 3001-particle cloud with 15 % static particles.  Stored per case: rows
X(3), Gamma(3), sigma, static = 8 x N.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from vpm_import import load  # noqa: E402


def cases():
    vpm = load()
    gold = np.load(os.path.join(HERE, "p2p_golden.npz"))
    n = gold["X"].shape[1]
    a = np.zeros((8, n))
    a[0:3], a[3:6], a[6], a[7] = gold["X"], gold["Gamma"], gold["sigma"], gold["static"]
    yield "gold40", a
    pf = vpm.fields.ring_field(Nphi=100, nc=3)
    yield "ring_c1", np.vstack([pf.particles[0:7, :pf.np], pf.particles[42:43, :pf.np]])
    pf = vpm.fields.cloud_field(3001, static_fraction=0.15, seed=9)
    yield "cloud3001", np.vstack([pf.particles[0:7, :pf.np], pf.particles[42:43, :pf.np]])


def write(path, arrays):
    off = 0
    with open(path + ".f64", "wb") as fb, open(path + ".txt", "w") as ft:
        for name, a in arrays:
            a = np.asfortranarray(a, dtype="<f8")
            fb.write(a.tobytes(order="F"))
            ft.write(f"{name} {a.shape[0]} {a.shape[1]} {off}\n")
            off += a.size


def read(path):
    """dict name -> array (rows x cols, Fortran order) of a .f64/.txt pair"""
    out = {}
    raw = np.fromfile(path + ".f64", dtype="<f8")
    with open(path + ".txt") as ft:
        for line in ft:
            name, r, c, off = line.split()
            r, c, off = int(r), int(c), int(off)
            out[name] = raw[off:off + r * c].reshape((r, c), order="F")
    return out


if __name__ == "__main__":
    write(os.path.join(HERE, "ref_inputs"), list(cases()))
    print({k: v.shape for k, v in read(os.path.join(HERE, "ref_inputs")).items()})
