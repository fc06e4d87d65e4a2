"""Device-built leaf lists (vpm_leaflists_build / vpm_uj_nearfield).

Index work: the device lists must equal the CPU restatement (oracle/leaflists.py) bit for bit.
Arithmetic: the near field evaluated over the resident lists must equal fmm.direct!'s arithmetic
over the same lists (oracle.direct_leafpairs) to 1e-12."""
import numpy as np
import pytest

from helpers import TOL_FP64, relerr
from oracle import leaflists, oracle

pytestmark = pytest.mark.gpu


def same_lists(a, b):
    return all(np.array_equal(a[k], b[k]) for k in ("sort_index", "leaf_begin", "leaf_end", "direct_list"))


@pytest.mark.parametrize("n,ncrit", [(1, 8), (2, 8), (257, 4), (5000, 16), (5000, 200), (40000, 64)])
def test_device_lists_equal_cpu_restatement_cloud(vpm, handle, n, ncrit):
    pf = vpm.fields.cloud_field(n, seed=31 + n)
    dev = vpm.leaf_lists(pf, ncrit=ncrit, theta=0.4)
    ref = leaflists.build_leaf_lists(pf.get_X(), pf.get_sigma(), ncrit=ncrit, theta=0.4)
    assert same_lists(dev, ref)
    assert sorted(dev["sort_index"].tolist()) == list(range(n))


@pytest.mark.parametrize("theta", [0.25, 0.4, 0.7])
def test_device_lists_ring_and_jet(vpm, handle, theta):
    """planar-ish (ring: thin z extent, padded) and elongated (jet) fields"""
    for pf in (vpm.fields.ring_field(Nphi=100, nc=3), vpm.fields.jet_field(n_target=6000)):
        dev = vpm.leaf_lists(pf, ncrit=32, theta=theta)
        ref = leaflists.build_leaf_lists(pf.get_X(), pf.get_sigma(), ncrit=32, theta=theta)
        assert same_lists(dev, ref)


def test_device_lists_degenerate_field(vpm, handle):
    """all particles at one point: one cell, one leaf, one self pair"""
    pf = vpm.fields.cloud_field(50, seed=1)
    pf.particles[0:3, :50] = np.array([[0.3], [0.1], [2.0]])
    dev = vpm.leaf_lists(pf, ncrit=8, theta=0.4)
    ref = leaflists.build_leaf_lists(pf.get_X(), pf.get_sigma(), ncrit=8, theta=0.4)
    assert same_lists(dev, ref)
    assert len(dev["leaf_begin"]) == 1 and dev["direct_list"].tolist() == [[0, 0]]


def nearfield_reference(pf, ll, kernel):
    order = ll["sort_index"]
    sb = np.asfortranarray(pf.particles[[0, 1, 2, 6, 3, 4, 5, 6]][:, :pf.np][:, order])
    tb = np.zeros((16, pf.np), order="F")
    tb[0:3] = pf.get_X()[:, order]
    leaves = (ll["leaf_begin"], ll["leaf_end"])
    oracle.direct_leafpairs(tb, sb, leaves, leaves, ll["direct_list"], kernel)
    out = np.zeros((12, pf.np))
    out[:, order] = tb[4:16]
    return out


@pytest.mark.parametrize("kernel", ["winckelmans", "gaussianerf", "singular", "gaussian"])
@pytest.mark.parametrize("ncrit", [20, 300])
def test_uj_nearfield_resident_lists(vpm, handle, kernel, ncrit):
    pf = vpm.fields.cloud_field(6000, kernel=vpm.KERNELS[kernel], static_fraction=0.1, seed=17)
    vpm.fields.random_results(pf, scale=1e-3)
    ll = vpm.leaf_lists(pf, ncrit=ncrit, theta=0.4)
    near = nearfield_reference(pf, ll, kernel)
    before = pf.particles.copy(order="F")
    # accumulate on what is there (e.g. a far field put there by the host)
    vpm.UJ_nearfield(pf, reset=False)
    assert relerr(pf.particles[9:12, :pf.np] - before[9:12, :pf.np], near[0:3]) < TOL_FP64
    assert relerr(pf.particles[15:24, :pf.np] - before[15:24, :pf.np], near[3:12]) < 1e-11  # difference of sums
    untouched = np.r_[0:9, 12:15, 24:46]
    assert np.array_equal(pf.particles[untouched], before[untouched])
    # reset=True: _reset_particles first (static particles keep and accumulate)
    pf.particles[:] = before
    vpm.UJ_nearfield(pf, reset=True)
    st = pf.get_static() != 0
    exp_U = np.where(st, before[9:12, :pf.np], 0.0) + near[0:3]
    exp_J = np.where(st, before[15:24, :pf.np], 0.0) + near[3:12]
    assert relerr(pf.particles[9:12, :pf.np], exp_U) < TOL_FP64 and relerr(pf.particles[15:24, :pf.np], exp_J) < TOL_FP64
    assert np.all(pf.particles[12:15, :pf.np][:, ~st] == 0) and np.all(pf.particles[24:27, :pf.np][:, ~st] == 0)


def test_uj_nearfield_whole_field_list_equals_direct(vpm, handle):
    """theta so small that every leaf pair is near field: the list form must reproduce UJ_direct"""
    pf = vpm.fields.cloud_field(3000, kernel=vpm.winckelmans, seed=4)
    ref = pf.particles.copy(order="F")
    oracle.uj_direct(ref, pf.np, "winckelmans", reset=True)
    ll = vpm.leaf_lists(pf, ncrit=64, theta=1e-3, fetch=True)
    nl = len(ll["leaf_begin"])
    assert len(ll["direct_list"]) == nl * nl
    vpm.UJ_nearfield(pf, reset=True)
    assert relerr(pf.particles[9:12], ref[9:12]) < TOL_FP64 and relerr(pf.particles[15:24], ref[15:24]) < TOL_FP64


def test_nearfield_requires_lists_and_bad_pairs_are_errors(vpm, handle):
    pf = vpm.fields.cloud_field(500, seed=2)
    vpm.leaf_lists(pf, ncrit=16)
    pf2 = vpm.fields.cloud_field(400, seed=2)
    with pytest.raises(vpm.VpmError) as e:
        vpm.UJ_nearfield(pf2)
    assert e.value.code == -6
    # Hook 3 with a list entry outside the leaf tables: EINVAL naming the entry
    ll = leaflists.build_leaf_lists(pf.get_X(), pf.get_sigma(), ncrit=16)
    dl = ll["direct_list"].copy()
    dl[7, 1] = len(ll["leaf_begin"]) + 3
    sb = np.asfortranarray(vpm.source_system_to_buffer(pf)[:, ll["sort_index"]])
    tb = np.zeros((16, pf.np), order="F")
    leaves = (ll["leaf_begin"], ll["leaf_end"])
    with pytest.raises(vpm.VpmError) as e:
        vpm.nearfield_device(tb, leaves, sb, leaves, dl, vpm.winckelmans)
    assert e.value.code == -1 and "pair 7" in str(e.value)


def test_leafpairs_run_to_run_bit_identical(vpm, handle):
    pf = vpm.fields.cloud_field(4000, kernel=vpm.gaussianerf, seed=8)
    ll = leaflists.build_leaf_lists(pf.get_X(), pf.get_sigma(), ncrit=24)
    rng = np.random.default_rng(0)
    dl = ll["direct_list"][rng.permutation(len(ll["direct_list"]))]   # unsorted: device radix sort path
    sb = np.asfortranarray(vpm.source_system_to_buffer(pf)[:, ll["sort_index"]])
    leaves = (ll["leaf_begin"], ll["leaf_end"])
    outs = []
    for _ in range(2):
        tb = np.zeros((16, pf.np), order="F")
        tb[0:3] = pf.get_X()[:, ll["sort_index"]]
        vpm.nearfield_device(tb, leaves, sb, leaves, dl, vpm.gaussianerf)
        outs.append(tb)
    assert np.array_equal(outs[0], outs[1])


def test_c5_nearfield_2p24_slice_parity(vpm, handle):
    """BASELINE config 5 at its size: 2^24 particles, leaf lists (ncrit 128, theta = 0.4) and the near field
    of UJ_fmm on the device; five target leaves (first, last, three inside) re-done by the CPU oracle over
    the same list (fmm.direct! arithmetic per (target leaf, source leaf) entry, list order)."""
    n = 1 << 24
    pf = vpm.fields.cloud_field(n, kernel=vpm.winckelmans)
    ll = vpm.leaf_lists(pf, ncrit=128, theta=0.4)
    lb, le, pt, ps, order = ll["leaf_begin"], ll["leaf_end"], ll["pair_tgt"], ll["pair_src"], ll["sort_index"]
    nl = len(lb)
    assert nl > 100_000 and len(pt) > 10_000_000
    assert np.all(pt[1:] >= pt[:-1])                       # grouped by target leaf
    assert le[-1] == n and np.array_equal(lb[1:], le[:-1])  # leaves tile the sorted order
    vpm.UJ_nearfield(pf, reset=True)
    tm = handle.timing()
    sizes = (le - lb).astype(np.int64)
    assert tm["uj_pairs"] == int((sizes[pt] * sizes[ps]).sum())
    starts = np.searchsorted(pt, np.arange(nl + 1))
    X, G, sg = pf.get_X(), pf.get_Gamma(), pf.get_sigma()
    worst = 0.0
    for leaf in (0, nl // 3, nl // 2, (2 * nl) // 3, nl - 1):
        tcols = order[lb[leaf]:le[leaf]]
        tb = np.zeros((16, len(tcols)), order="F")
        tb[0:3] = X[:, tcols]
        for sl in ps[starts[leaf]:starts[leaf + 1]]:
            scols = order[lb[sl]:le[sl]]
            sb = np.zeros((8, len(scols)), order="F")
            sb[0:3], sb[4:7], sb[3], sb[7] = X[:, scols], G[:, scols], sg[scols], sg[scols]
            oracle.direct_buffers(tb, 0, len(tcols), sb, 0, len(scols), "winckelmans")
        got = np.vstack([pf.particles[9:12, tcols], pf.particles[15:24, tcols]])
        worst = max(worst, relerr(got[0:3], tb[4:7]), relerr(got[3:12], tb[7:16]))
    assert worst < TOL_FP64, worst
    assert np.all(np.isfinite(pf.particles[9:24, :n]))


def test_nearfield_refuses_stale_lists(vpm, handle):
    """the resident lists are functions of X and sigma at build time: moving a particle (or changing a core size)
    without rebuilding must fail loudly, not drop near pairs silently; strengths may change freely"""
    pf = vpm.fields.cloud_field(3000, kernel=vpm.winckelmans, seed=8)
    vpm.leaf_lists(pf, ncrit=32, theta=0.4, fetch=False)
    vpm.UJ_nearfield(pf, reset=True)
    pf.particles[3:6, :pf.np] *= 2.0                    # strengths: fine
    vpm.UJ_nearfield(pf, reset=True)
    pf.particles[0, 17] += 1e-9                         # one particle moved by 1e-9
    with pytest.raises(vpm.VpmError, match="changed since vpm_leaflists_build"):
        vpm.UJ_nearfield(pf, reset=True)
    vpm.leaf_lists(pf, ncrit=32, theta=0.4, fetch=False)
    vpm.UJ_nearfield(pf, reset=True)
    pf.particles[6, 5] *= 1.0000001
    with pytest.raises(vpm.VpmError):
        vpm.UJ_nearfield(pf, reset=True)
