"""Builds the sm_100a shared library in-tree (gym_rotor_b200/csrc/libquadrotor_b200.so).

nvcc cross-compiles without a GPU.  The library has no torch dependency: plain nvcc -shared.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libquadrotor_b200.so")
SOURCES = ["quadrotor_b200.cu"]
DEPS = ["quadrotor_b200.cu", "qr_kernels.cuh", "qr_env.cuh", "qr_traj.cuh", "qr_dop853.cuh", "qr_math.cuh", "dop853_tableau.h",
        os.path.join("generated", "actor_td3.cuh"),
        os.path.join("..", "..", "include", "quadrotor_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False, out=None, defines=()):
    """out / defines: build an experimental variant next to the product library (tools/ab_build.py)."""
    if out is None and not force and not needs_build():
        return LIB
    extra = os.environ.get("QR_NVCC_EXTRA", "").split() if out else []   # experiments only (tools/ab_build.py), never the product build
    cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out or LIB] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("nvcc failed building %s" % LIB)
    if verbose:
        print(res.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
