"""High-precision arbiter (mpmath, 50 digits) of the reference FORMULAS -- TEST
INFRASTRUCTURE.  Not op-for-op: it evaluates the mathematics of
src/FLOWVPM_fmm.jl:113-161, src/FLOWVPM_kernel.jl:44-84 and
src/FLOWVPM_subfilterscale_models.jl:16-41 exactly (true erf, true exp), so it
bounds the rounding error of both the C oracle and the CUDA path.  Pure Python:
use on a handful of targets only."""
import mpmath as mp

mp.mp.dps = 50
_c4 = 1 / (4 * mp.pi)
_c2 = mp.sqrt(2 / mp.pi)
_c1 = 1 / (2 * mp.pi) ** mp.mpf("1.5")
_c3 = 3 / (4 * mp.pi)


def g_dgdr(kernel, s):
    s = mp.mpf(s)
    if kernel == "singular":
        return mp.mpf(1), mp.mpf(0)
    if kernel == "gaussianerf":
        aux = _c2 * s * mp.exp(-s * s / 2)
        return mp.erf(s / mp.sqrt(2)) - aux, s * aux
    if kernel == "gaussian":
        e = mp.exp(-s**3)
        return 1 - e, 3 * s * s * e
    if kernel == "winckelmans":
        a = (s * s + 1) ** mp.mpf("2.5")
        return s**3 * (s * s + mp.mpf("2.5")) / a, mp.mpf("7.5") * s * s / (a * (s * s + 1))
    raise ValueError(kernel)


def zeta(kernel, s):
    s = mp.mpf(s)
    if kernel == "singular":
        return mp.mpf(1) if s == 0 else mp.mpf(0)
    if kernel == "gaussianerf":
        return _c1 * mp.exp(-s * s / 2)
    if kernel == "gaussian":
        return _c3 * mp.exp(-s**3)
    if kernel == "winckelmans":
        return _c4 * mp.mpf("7.5") / (s * s + 1) ** mp.mpf("3.5")
    raise ValueError(kernel)


def uj_target(xt, X, Gamma, sigma, kernel):
    """U (3) and J (9, flat index i+3j = du_i/dx_j) induced on point xt by all sources."""
    U = [mp.mpf(0)] * 3
    J = [mp.mpf(0)] * 9
    xt = [mp.mpf(float(v)) for v in xt]
    n = X.shape[1]
    for s in range(n):
        d = [xt[k] - mp.mpf(float(X[k, s])) for k in range(3)]
        r2 = d[0] ** 2 + d[1] ** 2 + d[2] ** 2
        if r2 == 0:
            continue
        r = mp.sqrt(r2)
        G = [mp.mpf(float(Gamma[k, s])) for k in range(3)]
        sg = mp.mpf(float(sigma[s]))
        g, dg = g_dgdr(kernel, r / sg)
        r3inv = 1 / (r2 * r)
        crss = [-_c4 * r3inv * (d[1] * G[2] - d[2] * G[1]),
                -_c4 * r3inv * (d[2] * G[0] - d[0] * G[2]),
                -_c4 * r3inv * (d[0] * G[1] - d[1] * G[0])]
        for k in range(3):
            U[k] += g * crss[k]
        aux = dg / (sg * r) - 3 * g / r2
        aux2 = -_c4 * g * r3inv
        for j in range(3):
            for i in range(3):
                J[i + 3 * j] += aux * crss[i] * d[j]
        J[1] -= aux2 * G[2]
        J[2] += aux2 * G[1]
        J[3] += aux2 * G[2]
        J[5] -= aux2 * G[0]
        J[6] -= aux2 * G[1]
        J[7] += aux2 * G[0]
    return U, J


def sfs_target(it, X, Gamma, sigma, Jall, static, kernel, transposed=True):
    """SFS (3) of target particle `it` from all non-static sources (Estr_direct!)."""
    out = [mp.mpf(0)] * 3
    JT = [mp.mpf(float(v)) for v in Jall[:, it]]
    n = X.shape[1]
    for s in range(n):
        if static is not None and static[s]:
            continue
        d = [mp.mpf(float(X[k, s])) - mp.mpf(float(X[k, it])) for k in range(3)]
        r = mp.sqrt(d[0] ** 2 + d[1] ** 2 + d[2] ** 2)
        G = [mp.mpf(float(Gamma[k, s])) for k in range(3)]
        JS = [mp.mpf(float(v)) for v in Jall[:, s]]
        D = [JT[k] - JS[k] for k in range(9)]
        if transposed:
            S = [D[3 * k] * G[0] + D[3 * k + 1] * G[1] + D[3 * k + 2] * G[2] for k in range(3)]
        else:
            S = [D[k] * G[0] + D[k + 3] * G[1] + D[k + 6] * G[2] for k in range(3)]
        si = 1 / mp.mpf(float(sigma[s]))
        w = zeta(kernel, r * si) * si**3
        for k in range(3):
            out[k] += w * S[k]
    return out
