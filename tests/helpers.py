"""Shared test helpers: parity metric, kernel lists, oracle plumbing."""
import numpy as np

KERNELS = ["singular", "gaussian", "gaussianerf", "winckelmans"]

# Parity bar of BASELINE.json north_star: relative error <= 1e-12 in FP64,
# measured norm-wise per field (SURVEY section 7 "hard parts" (iii)):
#   max_i |a_i - b_i|_inf / max_i |b_i|_inf
TOL_FP64 = 1e-12
TOL_FP32 = 1e-5

U_ROWS, J_ROWS, SFS_ROWS = slice(9, 12), slice(15, 24), slice(39, 42)
W_ROWS, PSE_ROWS = slice(12, 15), slice(24, 27)


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.max(np.abs(b))
    if denom == 0:
        return float(np.max(np.abs(a)))
    return float(np.max(np.abs(a - b)) / denom)


def j_row_errors(P_test, P_ref, n):
    """Per-component error of the nine J rows, each against ITS OWN maximum (the device rebuilds
    J33 as -(J11 + J22), so row 9 must be looked at on its own); rows that vanish by symmetry
    (< 1e-3 of the largest J entry) are measured against that floor instead."""
    J, Jr = P_test[J_ROWS, :n], P_ref[J_ROWS, :n]
    top = np.max(np.abs(Jr)) if Jr.size else 0.0
    if top == 0:
        return float(np.max(np.abs(J))) if J.size else 0.0
    worst = 0.0
    for k in range(9):
        den = max(np.max(np.abs(Jr[k])), 1e-3 * top)
        worst = max(worst, float(np.max(np.abs(J[k] - Jr[k])) / den))
    return worst


def field_errors(P_test, P_ref, n, rows=("U", "J", "SFS")):
    sl = {"U": U_ROWS, "J": J_ROWS, "SFS": SFS_ROWS, "W": W_ROWS, "PSE": PSE_ROWS}
    out = {r: relerr(P_test[sl[r], :n], P_ref[sl[r], :n]) for r in rows}
    if "J" in rows:
        out["Jrow"] = j_row_errors(P_test, P_ref, n)
    return out


def assert_parity(P_test, P_ref, n, tol=TOL_FP64, rows=("U", "J", "SFS"), what=""):
    errs = field_errors(P_test, P_ref, n, rows)
    bad = {k: v for k, v in errs.items() if not (v <= tol)}
    assert not bad, f"parity {what}: {errs} exceeds {tol}"
    return errs


def stretching(P, n, transposed=True):
    """S = (Gamma . grad') U from J and the particle's own Gamma
    (src/FLOWVPM_timeintegration.jl:488-499)."""
    J = P[J_ROWS, :n]
    G = P[3:6, :n]
    if transposed:
        return np.stack([J[3 * k] * G[0] + J[3 * k + 1] * G[1] + J[3 * k + 2] * G[2] for k in range(3)])
    return np.stack([J[k] * G[0] + J[k + 3] * G[1] + J[k + 6] * G[2] for k in range(3)])
