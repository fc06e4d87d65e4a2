"""Build libvpm_cuda.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libvpm_cuda.so")
SOURCES = ["vpm_abi.cu"]
HEADERS = ["vpm_kernels.cuh", "vpm_kernels_f32.cuh", "vpm_kernels_tab.cuh", "vpm_tab_coeffs.cuh", "vpm_leaf.cuh", "vpm_leaf_f32.cuh", "vpm_csr.cuh", "vpm_tree.cuh", "vpm_math.cuh",
           "vpm_coeffs.cuh", "vpm_step.cuh",
           # host side: parts of the single translation unit vpm_abi.cu
           "vpm_host_base.cuh", "vpm_host_sweeps.cuh", "vpm_host_hook1.cuh", "vpm_host_multi.cuh",
           "vpm_host_lists.cuh", "vpm_host_field.cuh", "vpm_abi_core.cuh", "vpm_abi_lists.cuh",
           "vpm_abi_field.cuh", "vpm_abi_instr.cuh",
           os.path.join("..", "..", "include", "vpm_cuda.h")]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libvpm_cuda.so cannot be built (there is no CPU fallback)")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile csrc/*.cu -> csrc/libvpm_cuda.so; returns the library path."""
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", LIB] + SOURCES + ["-ldl", "-lpthread"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    # the image exports CC/CXX wrappers that lack libgomp specs; nvcc only needs a host g++
    env = dict(os.environ)
    res = subprocess.run(cmd, cwd=CSRC, env=env, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
