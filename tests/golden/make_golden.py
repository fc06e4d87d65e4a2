#!/usr/bin/env python
"""Generate tests/golden/p2p_golden.npz.

The reference (Julia + FastMultipole.jl) cannot run in this environment and ships
no golden vectors for this path, so the committed vectors are produced by the
50-digit mpmath evaluation of the reference FORMULAS (oracle/hp_oracle.py,
citing src/FLOWVPM_fmm.jl:113-161, src/FLOWVPM_kernel.jl:44-84,
src/FLOWVPM_subfilterscale_models.jl:16-41), rounded to FP64.  They are
independent of both the C oracle and the CUDA code.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hp_oracle  # noqa: E402

KERNELS = ["singular", "gaussian", "gaussianerf", "winckelmans"]


def main():
    rng = np.random.Generator(np.random.PCG64(12345))
    n = 40
    # jittered 5 x 4 x 2 lattice, spacing 0.1, sigma ~ 0.13: s = r/sigma spans ~0.4..6
    ix, iy, iz = np.meshgrid(np.arange(5), np.arange(4), np.arange(2), indexing="ij")
    X = np.stack([ix.ravel(), iy.ravel(), iz.ravel()]).astype(float) * 0.1
    X += (rng.random((3, n)) - 0.5) * 0.05
    Gamma = (rng.random((3, n)) - 0.5) * 0.2
    sigma = 0.13 * (1 + (rng.random(n) - 0.5) * 0.2)
    static = np.zeros(n)
    static[[3, 17, 29]] = 1.0
    Jin = rng.standard_normal((9, n))  # the "final J" the SFS sweep reads
    out = dict(X=X, Gamma=Gamma, sigma=sigma, static=static, Jin=Jin)
    for k in KERNELS:
        U = np.zeros((3, n))
        J = np.zeros((9, n))
        for t in range(n):
            u, j = hp_oracle.uj_target(X[:, t], X, Gamma, sigma, k)
            U[:, t] = [float(v) for v in u]
            J[:, t] = [float(v) for v in j]
        out[f"U_{k}"] = U
        out[f"J_{k}"] = J
        for transposed in (True, False):
            S = np.zeros((3, n))
            for t in range(n):
                if static[t]:
                    continue
                S[:, t] = [float(v) for v in hp_oracle.sfs_target(t, X, Gamma, sigma, Jin, static, k, transposed)]
            out[f"SFS_{k}_{'T' if transposed else 'C'}"] = S
        print("done", k)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "p2p_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
