"""Build the sm_100a shared library in-tree (lfit_python_b200/liblfit_b200.so).

nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the
repo snapshot (it is git-ignored, not gpurun-ignored).
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "lfit_cabi.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "cv_kernels.cuh"), os.path.join(HERE, "csrc", "roche_device.cuh"),
        os.path.join(HERE, "csrc", "gp_device.cuh"), os.path.join(HERE, "csrc", "sampler.cuh"),
        os.path.join(HERE, "csrc", "angle_table.inc"), os.path.join(HERE, "csrc", "peer.cuh"),
        os.path.join(HERE, "..", "include", "lfit_b200.h")]
LIB = os.path.join(HERE, "liblfit_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    """Compile csrc/lfit_kernels.cu for sm_100a; returns the library path."""
    if not force and not stale():
        return LIB
    extra = os.environ.get("LFB_NVCC_EXTRA", "").split()     # tuning experiments, e.g. -DLFB_ELEM_BLOCKS=8
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, SRC]
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
