"""Reference-run golden vectors (tests/golden/ref_outputs.*, produced by baseline/julia/make_golden.jl
from the true FLOWVPM.jl): the oracle (CPU, `-m "not gpu"`) and the CUDA path (`-m gpu`) against what
the reference itself computed.  The reference cannot run in this repository's containers, so the
outputs file may be absent: then these tests are SKIPPED and parity stays pinned to mpmath / analytic
answers only.  The inputs file and this loader are committed so that a Julia owner can close the loop
with one command (see tests/golden/make_ref_inputs.py)."""
import os
import sys

import numpy as np
import pytest

from helpers import KERNELS, TOL_FP64, relerr
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_ref_inputs  # noqa: E402

INP = os.path.join(HERE, "golden", "ref_inputs")
OUT = os.path.join(HERE, "golden", "ref_outputs")
have_outputs = os.path.exists(OUT + ".f64") and os.path.exists(OUT + ".txt")
needs_reference = pytest.mark.skipif(not have_outputs, reason="tests/golden/ref_outputs.* absent: run "
                                     "baseline/julia/make_golden.jl with Julia + FLOWVPM.jl to pin parity to the reference")
ROWS = list(range(9, 12)) + list(range(15, 24)) + list(range(39, 42))
CASES = ["gold40", "ring_c1", "cloud3001"]


def field_of(vpm, a, kernel, transposed=True):
    n = a.shape[1]
    pf = vpm.ParticleField(n, kernel=vpm.KERNELS[kernel], transposed=transposed)
    pf.particles[0:7, :n] = a[0:7]
    pf.particles[42, :n] = a[7]
    pf.np = n
    return pf


def test_inputs_file_is_current(vpm):
    """the committed inputs are exactly what make_ref_inputs.py generates today (CPU, always runs)"""
    stored = make_ref_inputs.read(INP)
    for name, a in make_ref_inputs.cases():
        assert np.array_equal(stored[name], a), name
    assert sorted(stored) == sorted(CASES)


def compare(P, ref, what):
    got = P[ROWS]
    for rows, name in ((slice(0, 3), "U"), (slice(3, 12), "J"), (slice(12, 15), "SFS")):
        if np.abs(ref[rows]).max() == 0:
            assert np.abs(got[rows]).max() == 0, (what, name)
        else:
            assert relerr(got[rows], ref[rows]) < TOL_FP64, (what, name, relerr(got[rows], ref[rows]))


@needs_reference
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("kernel", KERNELS)
def test_oracle_matches_reference_run(vpm, case, kernel):
    inputs, outputs = make_ref_inputs.read(INP), make_ref_inputs.read(OUT)
    for tag, transposed in (("T", True), ("C", False)):
        pf = field_of(vpm, inputs[case], kernel, transposed)
        oracle.uj_direct(pf.particles, pf.np, kernel, sfs=True, reset=True, reset_sfs=True, transposed=transposed)
        compare(pf.particles[:, :pf.np], outputs[f"{case}/{kernel}/{tag}"], f"oracle {case}/{kernel}/{tag}")
        if transposed:
            oracle.uj_direct(pf.particles, pf.np, kernel, sfs=False, reset=False, reset_sfs=False)
            compare(pf.particles[:, :pf.np], outputs[f"{case}/{kernel}/accumulate"], f"oracle {case}/{kernel}/accumulate")


@needs_reference
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("kernel", KERNELS)
def test_cuda_matches_reference_run(vpm, handle, case, kernel):
    inputs, outputs = make_ref_inputs.read(INP), make_ref_inputs.read(OUT)
    for tag, transposed in (("T", True), ("C", False)):
        pf = field_of(vpm, inputs[case], kernel, transposed)
        vpm.UJ_direct(pf, sfs=True, reset=True, reset_sfs=True)
        compare(pf.particles[:, :pf.np], outputs[f"{case}/{kernel}/{tag}"], f"cuda {case}/{kernel}/{tag}")
        if transposed:
            vpm.UJ_direct(pf, sfs=False, reset=False, reset_sfs=False)
            compare(pf.particles[:, :pf.np], outputs[f"{case}/{kernel}/accumulate"], f"cuda {case}/{kernel}/accumulate")


def test_loader_roundtrip_with_oracle_outputs(vpm, tmp_path):
    """the loader path end to end without Julia: write what the ORACLE computes in make_golden.jl's output
    format, read it back, compare -- proves the file format, the row selection and the comparison code"""
    inputs = make_ref_inputs.read(INP)
    arrays = []
    for kernel in ("winckelmans", "gaussianerf"):
        pf = field_of(vpm, inputs["gold40"], kernel)
        oracle.uj_direct(pf.particles, pf.np, kernel, sfs=True, reset=True, reset_sfs=True)
        arrays.append((f"gold40/{kernel}/T", pf.particles[ROWS, :pf.np].copy()))
    make_ref_inputs.write(str(tmp_path / "ref_outputs"), arrays)
    back = make_ref_inputs.read(str(tmp_path / "ref_outputs"))
    for name, a in arrays:
        assert np.array_equal(back[name], a)
        pf = field_of(vpm, inputs["gold40"], name.split("/")[1])
        oracle.uj_direct(pf.particles, pf.np, name.split("/")[1], sfs=True, reset=True, reset_sfs=True)
        compare(pf.particles[:, :pf.np], back[name], name)
