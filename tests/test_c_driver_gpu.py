"""The C ABI driven from plain C (tests/c_abi_driver.c, compiled with gcc against
include/vpm_cuda.h and linked to libvpm_cuda.so) -- no Python, no torch on the product side."""
import os
import subprocess

import numpy as np
import pytest

from helpers import TOL_FP64, relerr
from oracle import leaflists, oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("kernel_id,kernel", [(3, "winckelmans"), (2, "gaussianerf")])
def test_plain_c_driver(tmp_path, kernel_id, kernel):
    exe = str(tmp_path / "c_abi_driver")
    libdir = os.path.join(ROOT, "flowvpm.jl_b200", "csrc")
    subprocess.run(["/usr/bin/gcc", "-std=c11", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_abi_driver.c"), "-L", libdir, "-lvpm_cuda", "-lm",
                    f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    n = 300
    out = subprocess.run([exe, str(n), str(kernel_id)], check=True, capture_output=True, text=True).stdout.splitlines()
    assert out[0].startswith("abi 1 gpus 1")
    got = np.array([[float(v) for v in line.split()] for line in out[1:]]).T   # 15 x n
    i = np.arange(n)
    t = 0.05 * i
    P = np.zeros((46, n), order="F")
    P[0], P[1], P[2] = np.cos(t), np.sin(t), 0.02 * t
    P[3], P[4], P[5] = -0.01 * np.sin(t), 0.01 * np.cos(t), 0.001
    P[6] = 0.08 + 0.01 * np.sin(3 * t)
    P[42] = (i % 17 == 0).astype(float)
    P[9] = 0.5
    oracle.uj_direct(P, n, kernel, sfs=True, reset=True, reset_sfs=True)
    assert relerr(got[0:3], P[9:12]) < TOL_FP64
    assert relerr(got[3:12], P[15:24]) < TOL_FP64
    assert relerr(got[12:15], P[39:42]) < TOL_FP64


def test_plain_c_driver_leaf_lists_and_nearfield(tmp_path):
    """the f-3 entry points from plain C: device-built lists (counts equal to the CPU restatement)
    and the near field over them (equal to fmm.direct! over the restated lists)"""
    exe = str(tmp_path / "c_abi_driver")
    libdir = os.path.join(ROOT, "flowvpm.jl_b200", "csrc")
    subprocess.run(["/usr/bin/gcc", "-std=c11", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_abi_driver.c"), "-L", libdir, "-lvpm_cuda", "-lm",
                    f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    n = 900
    out = subprocess.run([exe, str(n), "3", "1"], check=True, capture_output=True, text=True).stdout.splitlines()
    nl, npairs, covered = (int(v) for v in out[0].split()[1:])
    got = np.array([[float(v) for v in line.split()] for line in out[2:]]).T
    i = np.arange(n)
    t = 0.05 * i
    P = np.zeros((46, n), order="F")
    P[0], P[1], P[2] = np.cos(t), np.sin(t), 0.02 * t
    P[3], P[4], P[5] = -0.01 * np.sin(t), 0.01 * np.cos(t), 0.001
    P[6] = 0.08 + 0.01 * np.sin(3 * t)
    ll = leaflists.build_leaf_lists(P[0:3], P[6], ncrit=16, theta=0.4)
    assert (nl, npairs, covered) == (len(ll["leaf_begin"]), len(ll["direct_list"]), n)
    order = ll["sort_index"]
    sb = np.asfortranarray(P[[0, 1, 2, 6, 3, 4, 5, 6]][:, order])
    tb = np.zeros((16, n), order="F")
    tb[0:3] = P[0:3, order]
    leaves = (ll["leaf_begin"], ll["leaf_end"])
    oracle.direct_leafpairs(tb, sb, leaves, leaves, ll["direct_list"], "winckelmans")
    near = np.zeros((12, n))
    near[:, order] = tb[4:16]
    st = (i % 17 == 0)
    exp_U = near[0:3].copy()
    exp_U[0, st] += 0.5   # static particles are not reset: the stale U stays under the sum
    assert relerr(got[0:3], exp_U) < TOL_FP64
    assert relerr(got[3:12], near[3:12]) < TOL_FP64
