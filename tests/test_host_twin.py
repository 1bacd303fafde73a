"""The arithmetic of the step kernel, checked WITHOUT a GPU.

tests/host_twin/twin.cpp compiles the per-env device functions of gym_rotor_b200/csrc (qr_math.cuh, qr_dop853.cuh,
qr_env.cuh, qr_traj.cuh -- the very headers nvcc compiles into the kernels) for the host with g++, float64
instantiations, and strings them together the way qr::k_step does for one lane.  Here that twin is run against the
reference's golden vectors with the tolerances of the -m gpu tests.  This is test infrastructure: it guards the
device code against logic errors on machines without a B200; the warp-level machinery of the kernel (lane refill,
stash, reset queue, stores) and the float32 paths are covered by tests/test_gpu_parity.py only.  The package never
loads the twin -- the product has no CPU path (tests/test_cabi_symbols.py enforces that).
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
TWIN_DIR = os.path.join(ROOT, "tests", "host_twin")
CUDA_INC = "/usr/local/cuda/include"

pytestmark = pytest.mark.skipif(shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")),
                                reason="needs g++ and the CUDA headers")


@pytest.fixture(scope="module")
def twin(tmp_path_factory):
    out = os.path.join(str(tmp_path_factory.mktemp("twin")), "libtwin.so")
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I" + CUDA_INC, "-I" + TWIN_DIR,
           "-I" + os.path.join(ROOT, "gym_rotor_b200", "csrc"), "-o", out, os.path.join(TWIN_DIR, "twin.cpp")]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    L = C.CDLL(out)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.tw_step.argtypes = [C.c_void_p, dp, dp, dp, dp, dp, C.c_int, dp, dp, C.POINTER(C.c_float), dp, ip, ip, ip]
    L.tw_step_f32.argtypes = L.tw_step.argtypes
    L.tw_reset.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, dp, dp, dp, dp]
    L.tw_traj_start.argtypes = [dp, dp, dp]
    L.tw_traj_desired.argtypes = [C.c_int, dp, dp, dp, dp, dp, dp, C.c_double, C.c_double, C.c_double]
    return L


def _config(mode, **kw):
    """qr_default_config of the real library (pure host code: no device needed)."""
    from gym_rotor_b200 import _native as nat
    cfg = nat.QrConfig()
    nat.check(nat.load().qr_default_config(C.byref(cfg), mode, nat.F64))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _step_rows(L, cfg, state, integ, params, goal, action, a32, f32=False):
    n = state.shape[0]
    step = L.tw_step_f32 if f32 else L.tw_step
    O = 23 if cfg.mode == 1 else 18
    st = np.empty((n, 18)); ig = np.empty((n, 8)); obs = np.empty((n, O), np.float32)
    rew = np.empty((n, 2)); dn = np.empty((n, 2), np.int32); nfev = np.empty(n, np.int32); nproj = np.empty(n, np.int32)
    status = np.empty(n, np.int32)
    goal = np.ascontiguousarray(goal, dtype=np.float64).copy()
    for i in range(n):
        s_i, g_i = np.ascontiguousarray(state[i], np.float64), np.ascontiguousarray(integ[i], np.float64)
        p_i, a_i = np.ascontiguousarray(params[i], np.float64), np.ascontiguousarray(action[i], np.float64)
        status[i] = step(C.byref(cfg), _dp(s_i), _dp(g_i), _dp(p_i), _dp(goal[i]), _dp(a_i), int(a32),
                              _dp(st[i]), _dp(ig[i]), obs[i].ctypes.data_as(C.POINTER(C.c_float)), _dp(rew[i]),
                              dn[i].ctypes.data_as(C.POINTER(C.c_int)), nfev[i:i + 1].ctypes.data_as(C.POINTER(C.c_int)),
                              nproj[i:i + 1].ctypes.data_as(C.POINTER(C.c_int)))
    return st, ig, obs, rew, dn, nfev, status, goal


def _relerr(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


@pytest.mark.parametrize("fw,tag,a32", [("MONO", "mono", False), ("MONO", "mono", True),
                                        ("MODUL", "modul", False), ("MODUL", "modul", True)])
def test_device_functions_match_reference_golden(twin, fw, tag, a32):
    """Same vectors and tolerances as test_gpu_parity.py::test_fp64_step_matches_reference_golden."""
    g = np.load(os.path.join(G, "step_%s_%s.npz" % (tag, "a32" if a32 else "a64")))
    cfg = _config(1 if fw == "MONO" else 2)
    act = g["action"].astype(np.float32).astype(np.float64) if a32 else g["action"]
    st, ig, obs, rew, dn, nfev, status, _ = _step_rows(twin, cfg, g["state_in"], g["integ_in"], g["params"], g["goal"], act, a32)
    assert _relerr(st, g["state_out"]) <= 1e-12
    assert np.abs(ig - g["integ_out"]).max() <= 1e-12
    flips = int((obs.view(np.uint32) != g["obs"].view(np.uint32)).sum())
    assert flips <= 3, flips
    assert np.abs(obs - g["obs"]).max() <= 1.2e-7
    G_ = g["done"].shape[1]
    assert (dn[:, :G_].astype(bool) == g["done"]).all()
    assert (nfev == g["nfev"]).all()
    assert np.abs(rew[:, :G_] - g["reward"]).max() <= 2e-7 and (rew[:, :G_] != g["reward"]).mean() <= 5e-3
    assert int(status.max()) == 0


@pytest.mark.parametrize("fw,tag", [("MONO", "mono"), ("MODUL", "modul")])
def test_device_functions_float32_within_1e5_of_reference(twin, fw, tag):
    """The float32 instantiations (packed stage sums, reciprocal multiplies, float32 reward norms; exact 1/x and sqrt in
    place of the MUFU approximations) with the tolerances of test_gpu_parity.py::test_fp32_step_within_1e5_of_reference."""
    g = np.load(os.path.join(G, "step_%s_a64.npz" % tag))
    cfg = _config(1 if fw == "MONO" else 2)
    act = g["action"].astype(np.float32).astype(np.float64)
    st, ig, obs, rew, dn, nfev, status, _ = _step_rows(twin, cfg, g["state_in"], g["integ_in"], g["params"], g["goal"], act, True, f32=True)
    assert np.abs(st - g["state_out"]).max() <= 1e-5
    assert np.abs(ig - g["integ_out"]).max() <= 1e-5
    assert np.abs(obs - g["obs"]).max() <= 1e-5
    G_ = g["done"].shape[1]
    assert np.abs(rew[:, :G_] - g["reward"]).max() <= 1e-4
    assert (dn[:, :G_].astype(bool) != g["done"]).mean() <= 2e-3
    assert ((nfev - 2) // 12 != (g["nfev"] - 2) // 12).mean() <= 0.03
    assert int(status.max()) == 0


def test_device_goal_mode0_matches_reference_trajgen(twin):
    """Wd of trajectory mode 0, computed from the pre-step state as the kernel does (goal_mode = 1)."""
    g = np.load(os.path.join(G, "step_mono_a64.npz"))
    cfg = _config(1, goal_mode=1)
    goal_in = g["goal"].copy(); goal_in[:, 9:12] = 7.0
    st, ig, obs, rew, dn, nfev, status, goal = _step_rows(twin, cfg, g["state_in"], g["integ_in"], g["params"], goal_in, g["action"], False)
    assert np.abs(goal[:, 9:12] - g["goal"][:, 9:12]).max() <= 1e-12
    assert _relerr(st, g["state_out"]) <= 1e-12
    assert np.abs(obs - g["obs"]).max() <= 1.2e-7


@pytest.mark.parametrize("integ", ["solve_ivp", "euler"])
def test_device_quad_v0_base_env(twin, integ):
    g = np.load(os.path.join(G, "quad_v0.npz"))
    cfg = _config(0, integrator=1 if integ == "euler" else 0)
    s_in = g[integ + "_state_in"]
    st, ig, obs, rew, dn, nfev, status, _ = _step_rows(twin, cfg, s_in, np.zeros((s_in.shape[0], 8)), g[integ + "_params"],
                                                       g[integ + "_goal"], g[integ + "_action"], False)
    assert np.abs(st - g[integ + "_state_out"]).max() < 1e-12
    assert (dn[:, :1].astype(bool) == g[integ + "_done"]).all()
    assert np.abs(rew[:, :1] - g[integ + "_reward"]).max() < 1e-12


def test_device_free_running_episode_vs_c_oracle(twin):
    """A 300-step free run (state fed back, no re-sync) stays within 1e-9 of the C oracle; attempt counts equal."""
    import quad_oracle as qo
    orc = qo.COracle("MONO")
    rng = np.random.default_rng(4)
    st_o, ig_o, par = orc.reset_from_uniforms(rng.random((8, 20)))
    goal = np.zeros((8, 12)); goal[:, 6] = 1.0
    st_t, ig_t = st_o.copy(), ig_o.copy()
    cfg = _config(1)
    alive = np.ones(8, bool)
    for t in range(300):
        act = rng.uniform(-1, 1, (8, 4)) * 0.3
        obs_o, rew_o, done_o, nfev_o, _ = orc.step(st_o, ig_o, par, goal, act)
        st_t, ig_t, obs_t, rew_t, dn_t, nfev_t, status, _ = _step_rows(twin, cfg, st_t, ig_t, par, goal, act, False)
        assert (nfev_t[alive] == nfev_o[alive]).all(), t
        assert (dn_t[alive, 0].astype(bool) == np.asarray(done_o).reshape(8, -1)[alive, 0]).all()
        alive &= ~np.asarray(done_o).reshape(8, -1)[:, 0].astype(bool)
        if not alive.any():
            break
        assert _relerr(st_t[alive], st_o[alive]) <= 1e-9, t
    assert t > 20


def test_device_reset_matches_oracle(twin):
    """reset_env (Philox4x32-10 keyed by seed / global env id / episode) against the oracle's reset from the same draws."""
    import quad_oracle as qo
    seed, n = 77, 64
    cfg = _config(1, seed=seed)
    u = np.empty((n, 20))
    for i in range(n):
        for j in range(5):
            w = qo.philox4x32_10([i & 0xFFFFFFFF, i >> 32, 1, j], [seed & 0xFFFFFFFF, seed >> 32])
            u[i, 4 * j:4 * j + 4] = [(k + 0.5) * 2.0 ** -32 for k in w]
    for env_type, et in (("train", qo.ENV_TRAIN), ("eval", qo.ENV_EVAL)):
        st_o, ig_o, par_o = qo.COracle("MONO").reset_from_uniforms(u, et)
        st = np.empty((n, 18)); ig = np.empty((n, 8)); par = np.empty((n, 6)); gl = np.empty((n, 12))
        for i in range(n):
            twin.tw_reset(C.byref(cfg), i, 1, 0 if env_type == "train" else 1, _dp(st[i]), _dp(ig[i]), _dp(par[i]), _dp(gl[i]))
        assert np.abs(st - st_o).max() <= 1e-14 and np.abs(par - par_o).max() <= 1e-15 and (ig == 0).all()


@pytest.mark.parametrize("name,mode", [("hover", 1), ("circle", 5), ("eight", 6), ("circle_manual", 5), ("takeoff", 2), ("land", 3),
                                       ("land_low", 3), ("stay", 4)])
def test_device_trajectory_modes_match_reference(twin, name, mode):
    """traj_start / traj_desired call by call against the reference's TrajectoryGenerator (tests/golden/traj_modes.npz)."""
    g = np.load(os.path.join(G, "traj_modes.npz"))
    st, goal_ref, bdd_ref, t_ref = g[name + "_state"], g[name + "_goal"], g[name + "_b1d_dot"], g[name + "_t"]
    t_traj, w, smooth, theta0 = g[name + "_draws"]
    u_t, u_w = (t_traj - 2.0) / 3.0, (w + 0.15 * np.pi) / (0.3 * np.pi)
    ts = np.zeros(12); goal = np.zeros(12); goal[6] = 1.0
    s0 = np.ascontiguousarray(st[0], np.float64)
    twin.tw_traj_start(_dp(s0[0:3].copy()), _dp(s0[6:15].copy()), _dp(ts))
    worst = 0.0
    for i in range(len(t_ref)):
        s = np.ascontiguousarray(st[i], np.float64)
        x, v, R, W = s[0:3].copy(), s[3:6].copy(), s[6:15].copy(), s[15:18].copy()
        twin.tw_traj_desired(mode, _dp(x), _dp(v), _dp(R), _dp(W), _dp(ts), _dp(goal), u_t, u_w, 0.005)
        if name == "circle_manual" and int(ts[1]) & 1 and ts[8] > 2.0:
            ts[8] = 1.75            # the golden run shortened the circle (num_circles = 0) to reach manual mode
        worst = max(worst, np.abs(goal - goal_ref[i]).max(), np.abs(ts[9:11] - bdd_ref[i, 0:2]).max())
        assert abs(ts[0] - t_ref[i]) < 1e-12
    assert worst < 1e-11, worst
    assert bool(int(ts[1]) & 2) == bool(g[name + "_manual"][-1])
