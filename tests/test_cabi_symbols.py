"""CPU-only: the in-tree CUDA library loads and exports every symbol include/quadrotor_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "quadrotor_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qr_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    from gym_rotor_b200 import build, _native
    build.build()
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert sorted(_native.EXPORTS) == names, "python binding and header disagree"
    lib.qr_abi_version.restype = ctypes.c_int
    assert lib.qr_abi_version() == _native.ABI_VERSION == 2


def test_config_struct_layout_matches_header():
    """qr_default_config writes the reference's constants into the ctypes mirror of qr_config."""
    from gym_rotor_b200 import _native
    L = _native.load()
    c = _native.QrConfig()
    assert L.qr_default_config(ctypes.byref(c), _native.MODE_DECOUPLED, _native.F64) == 0
    assert (c.mode, c.dtype, c.n_envs, c.seed) == (2, 1, 1, 1992)
    assert abs(c.dt - 0.005) < 1e-18 and c.g == 9.81 and c.rtol == 1e-3 and c.atol == 1e-6
    assert (c.Cx, c.CIx, c.Cv, c.Cw12, c.Cb1, c.CIb1, c.CW3, c.CW) == (6.0, 0.1, 0.4, 0.6, 6.0, 0.1, 0.1, 0.6)
    assert (c.reward_min, c.reward_min_1, c.reward_min_2) == (-14.0, -8.0, -7.0)   # quad.py:81-88
    assert (c.alpha, c.beta, c.min_force, c.udm_pct, c.euler_lim_deg) == (0.01, 0.05, 0.5, 10.0, 85.0)
    assert L.qr_default_config(ctypes.byref(c), 7, 0) != 0
    assert b"bad arguments" in L.qr_last_error()


def test_no_device_fails_loudly():
    """Without a GPU the product must refuse to run -- there is no CPU fallback."""
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from gym_rotor_b200 import _native, vec_env
    with pytest.raises(_native.NativeError):
        vec_env.BatchedQuadEnv(8)
    L = _native.load()
    c = _native.QrConfig()
    L.qr_default_config(ctypes.byref(c), 1, 0)
    h = ctypes.c_void_p()
    rc = L.qr_create(ctypes.byref(c), 0, ctypes.byref(h))
    assert rc != 0 and b"no CUDA device" in L.qr_last_error()


def test_create_rejects_inconsistent_configs():
    """Argument checks of qr_create come before the device is touched: they can be exercised anywhere."""
    from gym_rotor_b200 import _native
    L = _native.load()
    h = ctypes.c_void_p()

    def rc_for(mode=1, **kw):
        c = _native.QrConfig()
        L.qr_default_config(ctypes.byref(c), mode, 0)
        c.n_envs = 64
        for k, v in kw.items():
            setattr(c, k, v)
        return L.qr_create(ctypes.byref(c), 0, ctypes.byref(h)), L.qr_last_error()
    rc, msg = rc_for(n_envs=0)
    assert rc == 1 and b"n_envs" in msg
    rc, msg = rc_for(goal_mode=_native.GOAL_TRAJ_STAY + 1)
    assert rc == 1 and b"goal_mode" in msg
    rc, msg = rc_for(goal_mode=_native.GOAL_TRAJ_TAKEOFF, autoreset=1)       # modes 2-4 are not in the kernel's reset path
    assert rc == 1 and b"autoreset" in msg
    rc, msg = rc_for(mode=0, goal_mode=_native.GOAL_TRAJ_MODE0)              # on-device goals need a wrapper mode
    assert rc == 1 and b"wrapper" in msg
    rc, msg = rc_for(mode=1, integrator=_native.INT_EULER)                   # Euler exists for Quad-v0 only (quad.py:252)
    assert rc == 1 and b"Euler" in msg


def test_step_kernel_sass_budget():
    """Static guard on the headline kernel (CoupledWrapper, float32, single-step launch, on-device goal): registers within
    the 12-warps-per-SM budget, and no local-memory traffic creeping into the persistent loop -- round 1 lost ~8 % to
    loop-carried values that a large out-of-line call pushed into local memory.  The only LDL/STL left
    belong to the rare re-projection call sites (a 9-word matrix handed over by pointer) and to the callees."""
    import re
    import shutil
    import subprocess
    from gym_rotor_b200 import _native
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    _native.load()
    name = "_ZN2qr6k_stepIfLi1ELb0ELb1ELb0EEEvNS_8StepArgsIT_EE"
    res = subprocess.run(["cuobjdump", "-res-usage", _native.LIB_PATH], capture_output=True, text=True).stdout
    m = re.search(re.escape(name) + r".*?\n\s*(REG:(\d+).*)", res, re.S)
    assert m, "kernel not found in the library"
    assert int(m.group(2)) <= 170, m.group(1)                 # 384 threads x 170 registers = one CTA per SM
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", name, _native.LIB_PATH], capture_output=True, text=True).stdout
    lines = [l for l in sass.split("\n") if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]
    assert 5000 < len(lines) < 9000, len(lines)
    local = [i for i, l in enumerate(lines) if re.search(r"\b(LDL|STL)\b", l)]
    # the persistent loop is the first ~3 300 instructions (the out-of-line routines follow it)
    in_loop = [i for i in local if i < 3300]
    assert len(in_loop) <= 60, (len(in_loop), in_loop[:10])     # 5 re-projection call sites x 10 words today
    # ... and all but a handful (ptxas spills ~8 loop-carried words at 168 registers) sit next to a CALL
    stray = [i for i in in_loop if "CALL" not in "".join(lines[max(0, i - 12):i + 12])]
    assert len(stray) <= 24, (len(stray), [lines[i] for i in stray[:4]])


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under gym_rotor_b200/ may reference it."""
    pkg = os.path.join(ROOT, "gym_rotor_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "quad_oracle" not in txt and "oracle/" not in txt.replace("under oracle/", ""), f
