"""Multi-GPU paths (need >= 2 GPUs; skipped otherwise):
  * one process driving G devices through one handle (the Julia caller's situation):
    targets block-sharded, NCCL all-gather of the final J before the SFS sweep;
  * one process per GPU (torch.distributed + NCCL) through flowvpm_jl_b200.sharding,
    the path bench.py --gpus N times."""
import os
import socket
import sys

import numpy as np
import pytest

from helpers import assert_parity, relerr, TOL_FP64
from oracle import leaflists, oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("kernel", ["winckelmans", "gaussianerf"])
@pytest.mark.parametrize("n", [5001, 64])
def test_single_process_multi_gpu_handle(vpm, kernel, n):
    g = _ngpu()
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    h = vpm.Handle(min(g, 4))
    try:
        pf = vpm.fields.cloud_field(n, kernel=vpm.KERNELS[kernel], static_fraction=0.1, seed=3)
        vpm.fields.random_results(pf, scale=1e-3)
        ref = pf.particles.copy(order="F")
        oracle.uj_direct(ref, n, kernel, sfs=True, reset=True, reset_sfs=True)
        vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True, handle=h)
        assert_parity(pf.particles, ref, n, rows=("U", "J", "SFS", "W", "PSE"))
        # accumulate semantics across devices
        oracle.uj_direct(ref, n, kernel, sfs=True, reset=False, reset_sfs=False)
        vpm.UJ_direct(pf, sfs=True, reset=False, reset_sfs=False, handle=h)
        assert_parity(pf.particles, ref, n)
        assert h.timing()["n_gpus"] == min(g, 4)
    finally:
        h.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from vpm_import import load
    vpm = load()
    from flowvpm_jl_b200 import sharding
    h = vpm.Handle(device_ids=[rank])
    pf = vpm.fields.cloud_field(n, kernel=vpm.gaussianerf, seed=44)
    src8 = vpm.source_system_to_buffer(pf)
    t0, t1 = sharding.shard_bounds(n, world, rank)
    local = torch.from_numpy(np.ascontiguousarray(src8[:, t0:t1].T)).cuda()
    f = sharding.ShardedField(h, local, n, rank, world, vpm.gaussianerf.id)
    uj = f.uj(0)
    sfs = f.sfs(8)
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, f"uj_{rank}.npy"), uj[: t1 - t0].cpu().numpy())
    np.save(os.path.join(out_dir, f"sfs_{rank}.npy"), sfs[: t1 - t0].cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_one_process_per_gpu_sharding(vpm, tmp_path):
    g = _ngpu()
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    n, world = 7001, 2
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    pf = vpm.fields.cloud_field(n, kernel=vpm.gaussianerf, seed=44)
    ref = pf.particles.copy(order="F")
    oracle.uj_direct(ref, n, "gaussianerf", sfs=True, reset=True, reset_sfs=True)
    uj = np.concatenate([np.load(tmp_path / f"uj_{r}.npy") for r in range(world)]).T
    sfs = np.concatenate([np.load(tmp_path / f"sfs_{r}.npy") for r in range(world)]).T
    assert relerr(uj[0:3], ref[9:12, :n]) < TOL_FP64
    assert relerr(uj[3:12], ref[15:24, :n]) < TOL_FP64
    assert relerr(sfs, ref[39:42, :n]) < TOL_FP64


@pytest.mark.parametrize("ncrit", [24, 300])
def test_nearfield_leafpairs_multi_gpu(vpm, ncrit):
    """Hook 3 with target leaves sharded over the devices of one handle"""
    g = _ngpu()
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    h = vpm.Handle(min(g, 4))
    try:
        pf = vpm.fields.cloud_field(6000, kernel=vpm.gaussianerf, seed=21)
        ll = leaflists.build_leaf_lists(pf.get_X(), pf.get_sigma(), ncrit=ncrit, theta=0.4)
        order = ll["sort_index"]
        sb = np.asfortranarray(vpm.source_system_to_buffer(pf)[:, order])
        rng = np.random.default_rng(1)
        tb = np.asfortranarray(rng.standard_normal((16, pf.np)) * 1e-3)   # accumulate on previous values
        tb[0:3] = pf.get_X()[:, order]
        leaves = (ll["leaf_begin"], ll["leaf_end"])
        ref = tb.copy(order="F")
        oracle.direct_leafpairs(ref, sb, leaves, leaves, ll["direct_list"], "gaussianerf")
        vpm.nearfield_device(tb, leaves, sb, leaves, ll["direct_list"], vpm.gaussianerf, handle=h)
        assert relerr(tb[4:7], ref[4:7]) < TOL_FP64 and relerr(tb[7:16], ref[7:16]) < TOL_FP64
        assert np.array_equal(tb[0:4], ref[0:4])
    finally:
        h.close()


@pytest.mark.parametrize("ncrit", [24, 300])
def test_uj_nearfield_device_lists_multi_gpu(vpm, ncrit):
    """f-3 on several devices: lists built on device 0, work items cut over the devices, sorted
    results returned to device 0 over NVLink; must equal the single-list oracle evaluation and
    the single-GPU result bit for bit (same per-target summation order)"""
    g = _ngpu()
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    hm, h1 = vpm.Handle(min(g, 4)), vpm.Handle(1)
    try:
        pf = vpm.fields.cloud_field(7000, kernel=vpm.winckelmans, static_fraction=0.05, seed=23)
        vpm.fields.random_results(pf, scale=1e-3)
        before = pf.particles.copy(order="F")
        ll = vpm.leaf_lists(pf, ncrit=ncrit, theta=0.4, handle=hm)
        ref_ll = leaflists.build_leaf_lists(pf.get_X(), pf.get_sigma(), ncrit=ncrit, theta=0.4)
        assert all(np.array_equal(ll[k], ref_ll[k]) for k in ref_ll)
        vpm.UJ_nearfield(pf, reset=False, handle=hm)
        multi = pf.particles.copy(order="F")
        pf.particles[:] = before
        vpm.leaf_lists(pf, ncrit=ncrit, theta=0.4, handle=h1, fetch=False)
        vpm.UJ_nearfield(pf, reset=False, handle=h1)
        assert np.array_equal(multi, pf.particles)
        order = ll["sort_index"]
        sb = np.asfortranarray(vpm.source_system_to_buffer(pf)[:, order])
        tb = np.zeros((16, pf.np), order="F")
        tb[0:3] = pf.get_X()[:, order]
        leaves = (ll["leaf_begin"], ll["leaf_end"])
        oracle.direct_leafpairs(tb, sb, leaves, leaves, ll["direct_list"], "winckelmans")
        near = np.zeros((12, pf.np))
        near[:, order] = tb[4:16]
        assert relerr(multi[9:12] - before[9:12], near[0:3]) < TOL_FP64
        assert relerr(multi[15:24] - before[15:24], near[3:12]) < 1e-11
    finally:
        hm.close()
        h1.close()


def test_leafpairs_level_ordered_branch_table_multi_gpu(vpm):
    """Hook 3 with a leaf table as a level-ordered tree hands it over (interior branches, leaves not
    in body order): the work items are laid out in body order on the device, so the devices still
    share the work (and overlapping WORK leaves fall back to one device); result = oracle"""
    from test_parity_gpu import branch_table_like_fastmultipole
    g = _ngpu()
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    h = vpm.Handle(min(g, 4))
    try:
        pf = vpm.fields.cloud_field(6000, kernel=vpm.gaussianerf, seed=33)
        ll = leaflists.build_leaf_lists(pf.get_X(), pf.get_sigma(), ncrit=30)
        lb, le, dl = branch_table_like_fastmultipole(ll, pf.np)
        order = ll["sort_index"]
        sb = np.asfortranarray(vpm.source_system_to_buffer(pf)[:, order])
        tb = np.zeros((16, pf.np), order="F")
        tb[0:3] = pf.get_X()[:, order]
        ref = tb.copy(order="F")
        oracle.direct_leafpairs(ref, sb, (lb, le), (lb, le), dl, "gaussianerf")
        vpm.nearfield_device(tb, (lb, le), sb, (lb, le), dl, vpm.gaussianerf, handle=h)
        assert relerr(tb[4:7], ref[4:7]) < TOL_FP64 and relerr(tb[7:16], ref[7:16]) < TOL_FP64
        # overlapping leaves that carry work: the root is also a target -> one device, still right
        dl2 = np.vstack([dl, [[0, len(lb) - 1]]]).astype(np.int32)
        tb[4:] = 0
        ref = tb.copy(order="F")
        oracle.direct_leafpairs(ref, sb, (lb, le), (lb, le), dl2, "gaussianerf")
        vpm.nearfield_device(tb, (lb, le), sb, (lb, le), dl2, vpm.gaussianerf, handle=h)
        assert relerr(tb[4:7], ref[4:7]) < TOL_FP64 and relerr(tb[7:16], ref[7:16]) < TOL_FP64
    finally:
        h.close()


def test_estr_and_zeta_leafpairs_multi_gpu(vpm):
    """Estr_fmm! / zeta_fmm over a list with the target leaves sharded over the devices: equal to
    the oracle and bit-identical to the single-device evaluation"""
    g = _ngpu()
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    hm, h1 = vpm.Handle(min(g, 4)), vpm.Handle(1)
    try:
        pf = vpm.fields.cloud_field(5000, kernel=vpm.gaussianerf, seed=29)
        vpm.fields.random_results(pf, scale=1e-2)
        ll = leaflists.build_leaf_lists(pf.get_X(), pf.get_sigma(), ncrit=40, theta=0.4)
        order, leaves = ll["sort_index"], (ll["leaf_begin"], ll["leaf_end"])
        dl = ll["direct_list"][::-1].copy()   # unsorted list
        before = pf.particles.copy(order="F")
        ref = before.copy(order="F")
        oracle.estr_leafpairs(ref, order, order, leaves, leaves, dl, "gaussianerf", True)
        vpm.Estr_fmm(pf, order, order, leaves, leaves, dl, handle=hm)
        multi = pf.particles.copy(order="F")
        pf.particles[:] = before
        vpm.Estr_fmm(pf, order, order, leaves, leaves, dl, handle=h1)
        assert np.array_equal(multi, pf.particles)
        assert relerr(multi[39:42], ref[39:42]) < TOL_FP64
        # zeta_fmm
        pf.particles[:] = before
        ref = before.copy(order="F")
        oracle.zeta_leafpairs(ref, order, leaves, dl, "gaussianerf")
        vpm.zeta_fmm(pf, order, leaves, dl, handle=hm)
        multi = pf.particles.copy(order="F")
        pf.particles[:] = before
        vpm.zeta_fmm(pf, order, leaves, dl, handle=h1)
        assert np.array_equal(multi, pf.particles)
        assert relerr(multi[15:18], ref[15:18]) < TOL_FP64
    finally:
        hm.close()
        h1.close()


@pytest.mark.parametrize("sfs", [False, "dynamic"])
def test_resident_step_multi_gpu(vpm, sfs):
    """vpm_field_step with the mirror replicated on every device of the handle: targets sharded,
    whole particle columns all-gathered after each sweep (NCCL), O(N) kernels run everywhere"""
    g = _ngpu()
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    h = vpm.Handle(min(g, 4))
    try:
        pf = vpm.fields.cloud_field(3001, kernel=vpm.gaussianerf, static_fraction=0.05, seed=51)
        ref = pf.particles.copy(order="F")
        kw = dict(integration="rungekutta3", f=0.0, g=0.2, sfs=sfs, clip_backscatter=bool(sfs), relaxation="pedrizzetti",
                  relax=True, rlxf=0.3, alpha=0.9, sfs_rlxf=0.3)
        rf = vpm.ResidentField(pf, handle=h)
        for _ in range(2):
            rf.nextstep(1e-3, **kw)
            oracle.field_step(ref, pf.np, "gaussianerf", 1e-3, transposed=True, **kw)
        rf.download()
        for rows in (slice(0, 7), slice(9, 12), slice(15, 24), slice(27, 42)):
            assert relerr(pf.particles[rows, :pf.np], ref[rows, :pf.np]) < 1e-9, rows
        # UJ_direct on the resident field across devices
        rf.UJ(sfs=True, reset=True, reset_sfs=True)
        oracle.uj_direct(ref, pf.np, "gaussianerf", sfs=True, reset=True, reset_sfs=True)
        rf.download()
        assert_parity(pf.particles, ref, pf.np, tol=1e-9)
    finally:
        h.close()


def test_corespreading_rbf_multi_gpu(vpm):
    """CoreSpreading + RBF conjugate gradient with the mirror replicated on several devices
    (zeta sweeps sharded + all-gathered, CG reductions read from device 0)"""
    g = _ngpu()
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    h = vpm.Handle(2)
    try:
        pf = vpm.fields.ring_field(Nphi=60, nc=1, kernel=vpm.gaussianerf, R=1.0, Rcross=0.15, sigma=0.12)
        pf.particles[7, :pf.np] = 4 / 3 * np.pi * 0.05**3
        ref = pf.particles.copy(order="F")
        vis = dict(nu=2e-3, sgm0=0.12, beta=1.02, itmax=20, tol=1e-4, iterror=True)
        vis_ref = dict(vis, t_sgm=0.0)
        kw = dict(integration="rungekutta3", f=0.0, g=0.2, sfs=False, relaxation="pedrizzetti", relax=True)
        rf = vpm.ResidentField(pf, handle=h)
        for _ in range(2):
            rf.nextstep(5e-2, viscous=vis, **kw)
            oracle.field_step(ref, pf.np, "gaussianerf", 5e-2, transposed=True, viscous=vis_ref, **kw)
        rf.download()
        for rows in (slice(0, 7), slice(9, 12), slice(15, 24), slice(27, 36)):
            assert relerr(pf.particles[rows, :pf.np], ref[rows, :pf.np]) < 1e-8, rows
    finally:
        h.close()


def test_corespreading_zeta_fmm_multi_gpu(vpm):
    """CoreSpreading(nu, sgm0, zeta_fmm) on a 2-GPU handle: the list sweep runs on device 0, its three sums per
    particle are broadcast and added on every mirror (which must stay identical: the next UJ sweep is sharded)"""
    g = _ngpu()
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    h = vpm.Handle(2)
    try:
        pf = vpm.fields.ring_field(Nphi=60, nc=1, kernel=vpm.gaussianerf, R=1.0, Rcross=0.15, sigma=0.12)
        pf.particles[7, :pf.np] = 4 / 3 * np.pi * 0.05**3
        ref = pf.particles.copy(order="F")
        vis = dict(nu=2e-3, sgm0=0.12, beta=1.02, itmax=20, tol=1e-4, iterror=True)
        vis_ref = dict(vis, t_sgm=0.0)
        kw = dict(integration="rungekutta3", f=0.0, g=0.2, sfs=False, relaxation="pedrizzetti", relax=True)
        rf = vpm.ResidentField(pf, handle=h)
        with oracle.cs_zeta_fmm(ncrit=20, theta=0.4, reset=True):
            for _ in range(3):
                rf.nextstep(5e-2, viscous=dict(vis, zeta="fmm_reset", ncrit=20, theta=0.4), **kw)
                oracle.field_step(ref, pf.np, "gaussianerf", 5e-2, transposed=True, viscous=vis_ref, **kw)
        rf.download()
        for rows in (slice(0, 7), slice(9, 12), slice(15, 24), slice(27, 36)):
            assert relerr(pf.particles[rows, :pf.np], ref[rows, :pf.np]) < 1e-8, rows
    finally:
        h.close()


def test_nearfield_device_call_shape_multi_gpu(vpm):
    """Hook 3 in the reference's call shape (vpm_nearfield_ranges) on a multi-GPU handle: target ranges cut
    over the devices, every device moving its own columns; bit-identical to one device"""
    g = _ngpu()
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    from test_hook3_shape_gpu import combine_source_indices, sorted_system, reference, SWITCH_UJ
    pf = vpm.fields.cloud_field(20000, kernel=vpm.winckelmans, seed=17)
    ll = leaflists.build_leaf_lists(pf.get_X(), pf.get_sigma(), ncrit=64, theta=0.4)
    ti, si = combine_source_indices(ll["direct_list"], ll["leaf_begin"], ll["leaf_end"])
    outs = []
    for ng in (1, min(g, 4)):
        h = vpm.Handle(ng)
        try:
            s = sorted_system(vpm, pf, ll["sort_index"])
            vpm.fmm_nearfield_device(s, ti, SWITCH_UJ, s, si, handle=h)
            outs.append(s.particles.copy(order="F"))
        finally:
            h.close()
    U, J = reference(sorted_system(vpm, pf, ll["sort_index"]), ll, "winckelmans")
    assert relerr(outs[0][9:12, :pf.np], U) < TOL_FP64 and relerr(outs[0][15:24, :pf.np], J) < TOL_FP64
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("ncrit", [24, 200])
def test_sharded_transfers_pinned_matrix_multi_gpu(vpm, ncrit):
    """page-locked matrix + reset + no static particles: every device pulls its 1/G of X, Gamma, sigma over its
    own PCIe link (all-gather over NVLink) for vpm_leaflists_build / vpm_uj_nearfield / vpm_nearfield_ranges and
    writes its 1/G of the result rows itself; results bit-identical to the single-GPU handle, untouched rows
    untouched"""
    g = _ngpu()
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    from test_hook3_shape_gpu import combine_source_indices, sorted_system, SWITCH_UJ
    hm, h1 = vpm.Handle(min(g, 4)), vpm.Handle(1)
    try:
        pf = vpm.fields.cloud_field(9001, kernel=vpm.winckelmans, seed=29)   # not a multiple of the device count
        vpm.fields.random_results(pf, scale=1e-3)
        before = pf.particles.copy(order="F")
        P = pf.particles
        hm.check(hm.lib.vpm_pin_host(hm.ptr, P.ctypes.data, P.nbytes))
        ll = vpm.leaf_lists(pf, ncrit=ncrit, theta=0.4, handle=hm)
        vpm.UJ_nearfield(pf, reset=True, handle=hm)
        multi = P.copy(order="F")
        hm.check(hm.lib.vpm_unpin_host(hm.ptr, P.ctypes.data))
        P[:] = before
        ll1 = vpm.leaf_lists(pf, ncrit=ncrit, theta=0.4, handle=h1)
        assert all(np.array_equal(ll[k], ll1[k]) for k in ("sort_index", "leaf_begin", "leaf_end", "direct_list"))
        vpm.UJ_nearfield(pf, reset=True, handle=h1)
        assert np.array_equal(multi, P)
        for rows in (slice(0, 9), slice(27, 46)):
            assert np.array_equal(P[rows], before[rows])
        # Hook 3 call shape with a page-locked source system
        ti, si = combine_source_indices(ll["direct_list"], ll["leaf_begin"], ll["leaf_end"])
        outs = []
        for h in (hm, h1):
            s = sorted_system(vpm, pf, ll["sort_index"])
            s.particles[9:27] = 0
            if h is hm:
                h.check(h.lib.vpm_pin_host(h.ptr, s.particles.ctypes.data, s.particles.nbytes))
            vpm.fmm_nearfield_device(s, ti, SWITCH_UJ, s, si, handle=h)
            if h is hm:
                h.check(h.lib.vpm_unpin_host(h.ptr, s.particles.ctypes.data))
            outs.append(s.particles.copy(order="F"))
        assert np.array_equal(outs[0], outs[1])
        order = ll["sort_index"]
        assert np.array_equal(outs[1][9:12, :pf.np], P[9:12, order]) and np.array_equal(outs[1][15:24, :pf.np], P[15:24, order])
    finally:
        hm.close()
        h1.close()
