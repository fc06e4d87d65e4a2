"""Hook 3 in the call shape the reference shows (src/FLOWVPM_gpu.jl:637-643):
    fmm.nearfield_device!(target_system, target_indices::Vector{UnitRange}, switch, source_system, source_indices)
Both systems are ParticleFields whose columns the ranges index; per target leaf the source ranges are
what combine_source_indices (:554-580) gathers from a direct_list sorted by target.  The host mirror
`vpm.fmm_nearfield_device` has the same argument list and calls vpm_nearfield_ranges; the Julia method
in flowvpm.jl_b200/julia/FLOWVPMCuda.jl does the same conversion."""
import numpy as np
import pytest

from helpers import TOL_FP64, relerr
from oracle import leaflists, oracle

pytestmark = pytest.mark.gpu

SWITCH_UJ = (False, True, True)  # DerivativesSwitch{PS,VS,GS}: UJ_fmm asks for velocity and its gradient


def sorted_system(vpm, pf, order):
    """the system in tree order (FastMultipole hands tree-sorted bodies to the device hook)"""
    s = vpm.ParticleField(pf.np, kernel=pf.kernel)
    s.particles[:, :pf.np] = pf.particles[:, order]
    s.np = pf.np
    return s


def combine_source_indices(direct_list, leaf_begin, leaf_end):
    """Python restatement of what the reference's helper produces: per target leaf (in order of first
    appearance in the target-sorted list) the list of source body ranges"""
    dl = np.asarray(direct_list)
    dl = dl[np.argsort(dl[:, 0], kind="stable")]
    targets, groups = [], []
    for t, s in dl:
        if not targets or targets[-1] != t:
            targets.append(int(t))
            groups.append([])
        groups[-1].append(range(int(leaf_begin[s]), int(leaf_end[s])))
    return [range(int(leaf_begin[t]), int(leaf_end[t])) for t in targets], groups


def reference(pf_sorted, ll, kernel, want_U=True, want_J=True):
    n = pf_sorted.np
    sb = np.asfortranarray(pf_sorted.particles[[0, 1, 2, 6, 3, 4, 5, 6]][:, :n])
    tb = np.zeros((16, n), order="F")
    tb[0:3] = pf_sorted.particles[0:3, :n]
    leaves = (ll["leaf_begin"], ll["leaf_end"])
    oracle.direct_leafpairs(tb, sb, leaves, leaves, ll["direct_list"], kernel, want_U, want_J)
    return tb[4:7], tb[7:16]


@pytest.mark.parametrize("kernel", ["winckelmans", "gaussianerf", "singular", "gaussian"])
@pytest.mark.parametrize("ncrit", [24, 200])
def test_nearfield_device_call_shape(vpm, handle, kernel, ncrit):
    pf = vpm.fields.cloud_field(6000, kernel=vpm.KERNELS[kernel], seed=13)
    ll = leaflists.build_leaf_lists(pf.get_X(), pf.get_sigma(), ncrit=ncrit, theta=0.4)
    sys_ = sorted_system(vpm, pf, ll["sort_index"])
    vpm.fields.random_results(sys_, scale=1e-3)           # the hook ACCUMULATES on what is there
    before = sys_.particles.copy(order="F")
    target_indices, source_indices = combine_source_indices(ll["direct_list"], ll["leaf_begin"], ll["leaf_end"])
    vpm.fmm_nearfield_device(sys_, target_indices, SWITCH_UJ, sys_, source_indices)
    U, J = reference(sys_, ll, kernel)
    n = sys_.np
    assert relerr(sys_.particles[9:12, :n] - before[9:12, :n], U) < TOL_FP64
    assert relerr(sys_.particles[15:24, :n] - before[15:24, :n], J) < 1e-11   # difference of O(1e-3) numbers
    # rows the hook must not touch: everything but U and J (vorticity rows 13:15 travel but are unchanged)
    for rows in (slice(0, 9), slice(12, 15), slice(24, 46)):
        assert np.array_equal(sys_.particles[rows], before[rows])


def test_nearfield_device_exact_accumulation_and_switches(vpm, handle):
    """from zero: results equal the list evaluation to 1e-12; VS only leaves J alone, GS only leaves U alone"""
    pf = vpm.fields.ring_field(Nphi=100, nc=3, kernel=vpm.gaussianerf)
    ll = leaflists.build_leaf_lists(pf.get_X(), pf.get_sigma(), ncrit=40, theta=0.4)
    sys_ = sorted_system(vpm, pf, ll["sort_index"])
    ti, si = combine_source_indices(ll["direct_list"], ll["leaf_begin"], ll["leaf_end"])
    U, J = reference(sys_, ll, "gaussianerf")
    n = sys_.np
    vpm.fmm_nearfield_device(sys_, ti, SWITCH_UJ, sys_, si)
    assert relerr(sys_.particles[9:12, :n], U) < TOL_FP64 and relerr(sys_.particles[15:24, :n], J) < TOL_FP64
    a = sys_.particles.copy(order="F")
    vpm.fmm_nearfield_device(sys_, ti, (False, True, False), sys_, si)
    assert np.array_equal(sys_.particles[15:24], a[15:24]) and relerr(sys_.particles[9:12, :n], 2 * U) < TOL_FP64
    vpm.fmm_nearfield_device(sys_, ti, (False, False, True), sys_, si)
    assert relerr(sys_.particles[9:12, :n], 2 * U) < TOL_FP64 and relerr(sys_.particles[15:24, :n], 2 * J) < TOL_FP64


def test_nearfield_device_warmup_shape_and_two_systems(vpm, handle):
    """the reference's own example of the call (warmup_gpu, src/FLOWVPM_gpu.jl:620-647): leaves that each
    span 1:n, a single source range per target leaf; and distinct source / target systems"""
    src = vpm.fields.cloud_field(700, kernel=vpm.winckelmans, seed=3)
    tgt = vpm.fields.cloud_field(450, kernel=vpm.winckelmans, seed=4)
    tgt.particles[0:3, :450] += 0.013
    ngpu = 2
    vpm.fmm_nearfield_device(tgt, [range(0, 450)] * ngpu, SWITCH_UJ, src, [range(0, 700)] * ngpu)
    sb = vpm.source_system_to_buffer(src)
    tb = np.zeros((16, 450), order="F")
    tb[0:3] = tgt.get_X()
    for _ in range(ngpu):
        oracle.direct_buffers(tb, 0, 450, sb, 0, 700, "winckelmans")
    assert relerr(tgt.get_U(), tb[4:7]) < TOL_FP64 and relerr(tgt.get_J(), tb[7:16]) < TOL_FP64
    assert np.all(src.particles[9:27] == 0)


def test_nearfield_device_edge_cases(vpm, handle):
    pf = vpm.fields.cloud_field(300, kernel=vpm.winckelmans, seed=8)
    base = pf.particles.copy(order="F")
    # a target leaf without sources, an empty target range, an empty source range: nothing happens
    vpm.fmm_nearfield_device(pf, [range(0, 100), range(100, 100), range(100, 300)], SWITCH_UJ, pf,
                             [[], [range(0, 300)], [range(50, 50)]])
    assert np.array_equal(pf.particles, base)
    vpm.fmm_nearfield_device(pf, [], SWITCH_UJ, pf, [])
    with pytest.raises(vpm.VpmError):
        vpm.fmm_nearfield_device(pf, [range(0, 400)], SWITCH_UJ, pf, [[range(0, 300)]])   # outside the field
    with pytest.raises(ValueError):
        vpm.fmm_nearfield_device(pf, [range(0, 10)], SWITCH_UJ, pf, [])
    with pytest.raises(vpm.VpmError):   # partially overlapping target ranges have no single owner per column
        vpm.fmm_nearfield_device(pf, [range(0, 100), range(50, 150)], SWITCH_UJ, pf, [[range(0, 300)], [range(0, 300)]])
    assert np.array_equal(pf.particles, base)
