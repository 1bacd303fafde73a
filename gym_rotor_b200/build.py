"""Builds the sm_100a shared library in-tree (gym_rotor_b200/csrc/libquadrotor_b200.so).

nvcc cross-compiles without a GPU.  The library has no torch dependency: plain nvcc, objects compiled in parallel
(the step kernel's instantiations are spread over ten translation units, csrc/qr_step_tu.cu), then one link.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libquadrotor_b200.so")
OBJ_DIR = os.path.join(CSRC, "_build")
DEPS = ["quadrotor_b200.cu", "qr_step_tu.cu", "qr_kernels.cuh", "qr_env.cuh", "qr_traj.cuh", "qr_dop853.cuh", "qr_math.cuh", "dop853_tableau.h",
        os.path.join("generated", "actor_td3.cuh"),
        os.path.join("..", "..", "include", "quadrotor_b200.h")]
# -prec-div / -prec-sqrt = false concern SINGLE precision only: the float32 mode's remaining `/` and sqrtf (its bar is 1e-5 of
# the reference) become MUFU + one Newton step without the slow-path call (+2 % at 128 steps per launch, -850 instructions of
# code; profiles/r02/r02ah_ab_fast_div.txt).  The float64 (parity) mode divides and takes roots in double precision, and its
# float32 observation / reward arithmetic uses explicit __f*_rn intrinsics, which these flags do not touch (GPU suite: 60 green).
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-prec-div=false", "-prec-sqrt=false",
              "-Xcompiler", "-fPIC"]
# (object name, source, defines): the C ABI + companion kernels, then the step kernel per (dtype, mode, policy)
UNITS = [("abi", "quadrotor_b200.cu", [])] + [
    ("step_%s_m%d_p%d" % (t, m, p), "qr_step_tu.cu", ["QR_TU_T=%s" % t, "QR_TU_MODE=%d" % m, "QR_TU_POLICY=%d" % p])
    for t in ("float", "double") for (m, p) in ((1, 0), (2, 0), (0, 0), (1, 1), (2, 1))]


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


def _compile(args):
    cmd, name = args
    res = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return name, res.returncode, res.stdout


def build(force=False, verbose=False, out=None, defines=(), only=None):
    """out / defines: build an experimental variant next to the product library (tools/ab_build.py).
    only: substrings of unit names to recompile (development; the other objects must exist already)."""
    if out is None and not force and not needs_build():
        return LIB
    extra = os.environ.get("QR_NVCC_EXTRA", "").split() if out else []   # experiments only (tools/ab_build.py), never the product build
    tag = hashlib.sha1((" ".join(sorted(defines)) + "|" + " ".join(extra)).encode()).hexdigest()[:10] if (defines or extra) else "product"
    odir = os.path.join(OBJ_DIR, tag)
    os.makedirs(odir, exist_ok=True)
    jobs, objs = [], []
    for name, src, defs in UNITS:
        obj = os.path.join(odir, name + ".o")
        objs.append(obj)
        if only is not None and os.path.exists(obj) and not any(o in name for o in only):
            continue
        # a wrapper source per unit, so that every cubin in the library has its own name (cuobjdump -xelf, tools/*.py)
        wrap = os.path.join(odir, name + ".cu")
        with open(wrap, "w") as f:
            f.write('#include "%s"\n' % os.path.join(CSRC, src))
        cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-I" + CSRC] + ["-D" + d for d in list(defs) + list(defines)] + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, wrap]
        jobs.append((cmd, name))
    workers = max(1, min(len(jobs), os.cpu_count() or 4))
    failed = False
    with concurrent.futures.ThreadPoolExecutor(workers) as ex:
        for name, rc, log in ex.map(_compile, jobs):
            if rc != 0:
                sys.stderr.write("---- %s\n%s" % (name, log))
                failed = True
            elif verbose:
                print("---- %s\n%s" % (name, log))
    if failed:
        raise RuntimeError("nvcc failed building %s" % (out or LIB))
    res = subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out or LIB] + objs, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("link failed for %s" % (out or LIB))
    return out or LIB


if __name__ == "__main__":
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
    print(build(force="--force" in sys.argv or bool(only), verbose="-v" in sys.argv, only=only or None))
